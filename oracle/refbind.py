"""oracle/refbind.py -- TEST INFRASTRUCTURE ONLY (imported by tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package).

ctypes bindings to the compiled, canonicalised reference in oracle/_ref/
(built by oracle/build_ref.sh from the unmodified sources under /root/reference).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
PIX_BYTES = 512 * 512 * 3

_enc = None
_dec = None


def available():
    return os.path.exists(os.path.join(_REF, "libnhwref_enc.so")) and os.path.exists(
        os.path.join(_REF, "libnhwref_dec.so"))


def enc_lib():
    global _enc
    if _enc is None:
        L = ctypes.CDLL(os.path.join(_REF, "libnhwref_enc.so"), mode=ctypes.RTLD_LOCAL)
        L.nhwref_encode.restype = ctypes.c_long
        L.nhwref_encode.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_long]
        L.nhwref_encode_discard.restype = ctypes.c_int
        L.nhwref_encode_discard.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.nhwref_encode_bmp_file.restype = ctypes.c_long
        L.nhwref_encode_bmp_file.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p]
        L.nhwref_tap_enable.argtypes = [ctypes.c_int]
        L.nhwref_tap_count.restype = ctypes.c_int
        L.nhwref_tap_name.restype = ctypes.c_char_p
        L.nhwref_tap_name.argtypes = [ctypes.c_int]
        L.nhwref_tap_data.restype = ctypes.c_void_p
        L.nhwref_tap_data.argtypes = [ctypes.c_int]
        L.nhwref_tap_bytes.restype = ctypes.c_size_t
        L.nhwref_tap_bytes.argtypes = [ctypes.c_int]
        L.nhwref_stage_colorspace.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p]
        L.nhwref_stage_pre_processing.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.nhwref_stage_dwt_y.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.nhwref_wavelet_analysis.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.nhwref_wavelet_synthesis.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int]
        _enc = L
    return _enc


_enc_stock = None


def enc_stock_lib():
    """The same reference objects on the stock allocator: CPU-baseline TIMING only."""
    global _enc_stock
    if _enc_stock is None:
        L = ctypes.CDLL(os.path.join(_REF, "libnhwref_enc_stock.so"), mode=ctypes.RTLD_LOCAL)
        L.nhwref_encode_discard.restype = ctypes.c_int
        L.nhwref_encode_discard.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _enc_stock = L
    return _enc_stock


def dec_lib():
    global _dec
    if _dec is None:
        L = ctypes.CDLL(os.path.join(_REF, "libnhwref_dec.so"), mode=ctypes.RTLD_LOCAL)
        L.nhwref_decode.restype = ctypes.c_long
        L.nhwref_decode.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_int]
        _dec = L
    return _dec


def _pix(a):
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
    assert a.size == PIX_BYTES, a.size
    return a


def ref_encode(pix, quality=20):
    """786432 raw BMP pixel bytes (file order) -> canonical reference .nhw bytes."""
    a = _pix(pix)
    out = np.zeros(1 << 20, dtype=np.uint8)
    n = enc_lib().nhwref_encode(a.ctypes.data, int(quality), out.ctypes.data, out.size)
    if n <= 0:
        raise RuntimeError("reference encoder failed: %d" % n)
    return out[:n].tobytes()


def ref_encode_taps(pix, quality=20):
    """-> (nhw bytes, {tap name: np.uint8 array})"""
    L = enc_lib()
    L.nhwref_tap_enable(1)
    try:
        data = ref_encode(pix, quality)
        taps = {}
        for i in range(L.nhwref_tap_count()):
            nb = L.nhwref_tap_bytes(i)
            buf = (ctypes.c_uint8 * nb).from_address(L.nhwref_tap_data(i))
            taps[L.nhwref_tap_name(i).decode()] = np.frombuffer(buf, dtype=np.uint8).copy()
    finally:
        L.nhwref_tap_enable(0)
    return data, taps


def ref_decode(nhw, planes=False):
    """.nhw bytes -> 786432 BMP pixel bytes (np.uint8); optionally also the Y/U/V planes."""
    src = np.frombuffer(bytes(nhw), dtype=np.uint8)
    out = np.zeros(PIX_BYTES, dtype=np.uint8)
    pl = np.zeros(3 * 262144, dtype=np.uint8)
    n = dec_lib().nhwref_decode(src.ctypes.data, src.size, out.ctypes.data, pl.ctypes.data, 1)
    if n != PIX_BYTES:
        raise RuntimeError("reference decoder failed: %d" % n)
    return (out, pl.reshape(3, 512, 512)) if planes else out


def ref_colorspace(pix, quality=20):
    a = _pix(pix)
    Y = np.zeros((512, 512), dtype=np.int16)
    U = np.zeros((256, 256), dtype=np.uint8)
    V = np.zeros((256, 256), dtype=np.uint8)
    enc_lib().nhwref_stage_colorspace(a.ctypes.data, int(quality), Y.ctypes.data, U.ctypes.data, V.ctypes.data)
    return Y, U, V


def ref_pre_processing(Y, quality=20):
    Y = np.ascontiguousarray(Y, dtype=np.int16).copy()
    enc_lib().nhwref_stage_pre_processing(Y.ctypes.data, int(quality))
    return Y


def ref_dwt_y(Y, quality=20):
    Y = np.ascontiguousarray(Y, dtype=np.int16)
    proc = np.zeros((512, 512), dtype=np.int16)
    ll1 = np.zeros((256, 256), dtype=np.int16)
    enc_lib().nhwref_stage_dwt_y(Y.ctypes.data, int(quality), proc.ctypes.data, ll1.ctypes.data)
    return proc, ll1


def ref_wavelet_analysis(jpeg, norder, last_stage, is_y, quality=20):
    """jpeg: int16 plane (512x512 for luma, 256x256 for chroma). -> (jpeg_after, proc_after)"""
    jpeg = np.ascontiguousarray(jpeg, dtype=np.int16).copy()
    proc = np.zeros_like(jpeg)
    enc_lib().nhwref_wavelet_analysis(jpeg.ctypes.data, proc.ctypes.data, norder, last_stage, is_y, quality)
    return jpeg, proc


def ref_wavelet_synthesis(jpeg, norder, last_stage, is_y):
    jpeg = np.ascontiguousarray(jpeg, dtype=np.int16).copy()
    proc = np.zeros_like(jpeg)
    enc_lib().nhwref_wavelet_synthesis(jpeg.ctypes.data, proc.ctypes.data, norder, last_stage, is_y)
    return jpeg, proc
