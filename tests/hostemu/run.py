"""tests/hostemu/run.py -- TEST TOOLING: run the host-compiled stage functions on one image
and report the first stage whose output differs from the reference tap of the same name."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refbind  # noqa: E402

_lib = None
TAPFN = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t)


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libhostemu.so")
        csrc = os.path.join(os.path.dirname(os.path.dirname(HERE)), "nhwcodec_b200", "csrc")
        srcs = [os.path.join(HERE, "hostemu.cpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
            import subprocess
            subprocess.check_call(["bash", os.path.join(HERE, "build.sh")])   # a stale library would test nothing
        _lib = ctypes.CDLL(so)
        _lib.he_new.restype = ctypes.c_void_p
        _lib.he_free.argtypes = [ctypes.c_void_p]
        _lib.he_encode.restype = ctypes.c_int
        _lib.he_encode.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p, TAPFN]
    return _lib


def host_encode(y_pre, u8, v8, q, want_taps=True):
    L = lib()
    h = L.he_new()
    taps = {}
    order = []

    def cb(name, ptr, n):
        buf = (ctypes.c_uint8 * n).from_address(ptr) if n else b""
        nm = name.decode()
        taps[nm] = np.frombuffer(buf, dtype=np.uint8).copy() if n else np.zeros(0, np.uint8)
        order.append(nm)

    out = np.zeros(1 << 20, dtype=np.uint8)
    y = np.ascontiguousarray(y_pre, dtype=np.int16)
    u = np.ascontiguousarray(u8, dtype=np.uint8)
    v = np.ascontiguousarray(v8, dtype=np.uint8)
    fn = TAPFN(cb) if want_taps else TAPFN(0)
    n = L.he_encode(h, y.ctypes.data, u.ctypes.data, v.ctypes.data, q, out.ctypes.data, fn)
    L.he_free(h)
    return (out[:n].tobytes() if n > 0 else n), taps, order


def host_decode(stream):
    """host-compiled decoder stage functions, run in the kernels' order -> 786432 pixel bytes"""
    L = lib()
    L.he_decode.restype = ctypes.c_int
    L.he_decode.argtypes = [ctypes.c_char_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p]
    rgb = np.zeros(786432, dtype=np.uint8)
    yuv = np.zeros(786432, dtype=np.uint8)
    # like the device copy of a chunk (api.cu), the stream is followed by 64 defined bytes: the bit reader looks a few
    # words past the last code
    rc = L.he_decode(bytes(stream) + b"\0" * 64, len(stream), rgb.ctypes.data, yuv.ctypes.data)
    return rc, rgb, yuv


def compare_decode(pix, q, verbose=True):
    stream = refbind.ref_encode(pix, q)
    want = refbind.ref_decode(stream)
    rc, rgb, _ = host_decode(stream)
    ok = rc == 0 and np.array_equal(rgb, np.asarray(want).reshape(-1))
    if verbose:
        print("decode:", "BIT-EXACT" if ok else "DIFFERENT (rc=%d)" % rc)
    return ok


def compare(pix, q, verbose=True):
    ref_stream, rt = refbind.ref_encode_taps(pix, q)
    y_pre = rt["y_pre"].view(np.int16)
    stream, ht, order = host_encode(y_pre, rt["cs_U"], rt["cs_V"], q)
    first_bad = None
    for name in order:
        if name not in rt:
            continue
        a, b = ht[name], rt[name]
        if name.endswith("_proc") or name.endswith("_jpeg") or name.endswith("_ll1"):
            a16, b16 = a.view(np.int16), b.view(np.int16)
            w = 512 if a16.size == 262144 else 256 if a16.size == 65536 else 128
            # chroma planes: only rows/cols that exist; luma tap planes compare fully except
            # regions the reference leaves as stale scratch (handled per-name below)
            if name.endswith("_jpeg"):
                # im_jpeg outside the band being reconstructed is stale scratch of earlier transforms in the reference
                # (nobody reads it): only the live N/2 x N/2 region is compared
                m = np.zeros((w, w), bool)
                m[: w // 2, : w // 2] = True
                a16 = np.where(m.reshape(-1), a16, 0)
                b16 = np.where(m.reshape(-1), b16, 0)
            same = np.array_equal(a16, b16)
            if not same:
                d = np.flatnonzero(a16 != b16)
                if verbose:
                    print("  %-14s MISMATCH %d cells, first at (r=%d,c=%d): got %d want %d" % (
                        name, d.size, d[0] // w, d[0] % w, a16[d[0]], b16[d[0]]))
                first_bad = first_bad or name
            elif verbose:
                print("  %-14s ok" % name)
        else:
            n = min(a.size, b.size)
            same = a.size == b.size and np.array_equal(a, b)
            if not same:
                d = np.flatnonzero(a[:n] != b[:n])
                if verbose:
                    print("  %-14s MISMATCH len %d vs %d, first diff at %s" % (name, a.size, b.size, d[:1]))
                first_bad = first_bad or name
            elif verbose:
                print("  %-14s ok" % name)
    ok = stream == ref_stream
    if verbose:
        print("stream:", "BIT-EXACT" if ok else "DIFFERENT", len(stream) if isinstance(stream, bytes) else stream,
              len(ref_stream))
    return ok, first_bad, stream, ref_stream


if __name__ == "__main__":
    from nhwcodec_b200 import synth
    kind = sys.argv[1] if len(sys.argv) > 1 else "natural"
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    q = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    pix = {"natural": synth.natural, "noise": synth.noise, "textured": synth.textured}[kind](seed)
    if os.environ.get("HE_DECODE"):
        compare_decode(pix, q)
    else:
        compare(pix, q)
