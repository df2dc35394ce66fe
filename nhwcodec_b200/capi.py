"""ctypes binding of include/nhw_cuda.h (the C-ABI of libnhw_cuda.so).

Mirrors the reference's per-image entry points (encoder/codec.h:184-189,
decoder/codec.h:198-202) as batch calls; see include/nhw_cuda.h for the contract.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnhw_cuda.so")

PIX_BYTES = 512 * 512 * 3
MAX_STREAM_BYTES = 1 << 19

# every symbol include/nhw_cuda.h declares (tests check the library exports all of them)
EXPORTS = [
    "nhw_create", "nhw_destroy", "nhw_last_error", "nhw_version", "nhw_encode_batch",
    "nhw_encode_batch_device", "nhw_decode_batch", "nhw_decode_batch_planes", "nhw_stage_frontend_device",
    "nhw_stage_colorspace_device", "nhw_synth_batch_device", "nhw_launch_count",
    "nhw_stream", "nhw_profile", "nhw_profile_read", "nhw_debug_stop_after", "nhw_debug_read", "nhw_debug_color_check",
    "nhw_decode_batch_device", "nhw_decode_batch_packed_device", "nhw_pack_batch_device", "nhw_digest_batch_device",
    "nhw_debug_dec_color_check", "nhw_host_alloc", "nhw_host_free",
]

_lib = None


class NhwError(RuntimeError):
    pass


def load_library():
    """Load libnhw_cuda.so; raises (loudly) if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NhwError("libnhw_cuda.so is not built (run `python -m nhwcodec_b200.build`); "
                       "this package has no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, u32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64
    L.nhw_create.argtypes = [i32, i32, ctypes.POINTER(vp)]
    L.nhw_create.restype = i32
    L.nhw_destroy.argtypes = [vp]
    L.nhw_destroy.restype = None
    L.nhw_last_error.restype = ctypes.c_char_p
    L.nhw_version.restype = i32
    L.nhw_encode_batch.argtypes = [vp, vp, i32, i32, vp, ctypes.c_size_t, vp, vp]
    L.nhw_encode_batch.restype = i32
    L.nhw_encode_batch_device.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    L.nhw_encode_batch_device.restype = i32
    L.nhw_decode_batch.argtypes = [vp, vp, vp, i32, vp, vp]
    L.nhw_decode_batch.restype = i32
    L.nhw_pack_batch_device.argtypes = [vp, vp, vp, i32, vp, vp]
    L.nhw_pack_batch_device.restype = i32
    L.nhw_digest_batch_device.argtypes = [vp, vp, ctypes.c_size_t, vp, u32, i32, vp]
    L.nhw_digest_batch_device.restype = i32
    L.nhw_decode_batch_device.argtypes = [vp, vp, ctypes.c_size_t, vp, i32, vp, vp]
    L.nhw_decode_batch_device.restype = i32
    L.nhw_decode_batch_packed_device.argtypes = [vp, vp, vp, i32, vp, vp]
    L.nhw_decode_batch_packed_device.restype = i32
    L.nhw_decode_batch_planes.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    L.nhw_decode_batch_planes.restype = i32
    L.nhw_stage_frontend_device.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp]
    L.nhw_stage_frontend_device.restype = i32
    L.nhw_stage_colorspace_device.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    L.nhw_stage_colorspace_device.restype = i32
    L.nhw_synth_batch_device.argtypes = [vp, vp, i32, u32, i32, vp]
    L.nhw_synth_batch_device.restype = i32
    L.nhw_launch_count.argtypes = [vp]
    L.nhw_launch_count.restype = u64
    L.nhw_stream.argtypes = [vp]
    L.nhw_stream.restype = vp
    L.nhw_debug_stop_after.argtypes = [vp, ctypes.c_char_p, i32]
    L.nhw_debug_stop_after.restype = i32
    L.nhw_debug_read.argtypes = [vp, ctypes.c_char_p, i32, vp, ctypes.c_size_t]
    L.nhw_debug_read.restype = i32
    L.nhw_debug_color_check.argtypes = [vp]
    L.nhw_debug_color_check.restype = ctypes.c_long
    L.nhw_debug_dec_color_check.argtypes = [vp]
    L.nhw_debug_dec_color_check.restype = ctypes.c_long
    L.nhw_profile.argtypes = [vp, i32]
    L.nhw_profile.restype = i32
    L.nhw_profile_read.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t]
    L.nhw_profile_read.restype = ctypes.c_long
    _lib = L
    return L


def _ptr(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class Codec:
    """One codec context on one GPU (one per process/rank).  Host-buffer calls mirror the
    reference's nhw-enc / nhw-dec per-image flow for a whole batch."""

    def __init__(self, device=0, max_batch=256):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.nhw_create(int(device), int(max_batch), ctypes.byref(h))
        if rc != 0:
            raise NhwError("nhw_create failed (%d): %s" % (rc, self.lib.nhw_last_error().decode()))
        self.h = h
        self.device = int(device)
        self.max_batch = int(max_batch)
        self._lut = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.nhw_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise NhwError("%s failed (%d): %s" % (what, rc, self.lib.nhw_last_error().decode()))

    @property
    def launches(self):
        return int(self.lib.nhw_launch_count(self.h))

    @property
    def stream_ptr(self):
        """cudaStream_t of this context (wrap with torch.cuda.ExternalStream to record events)"""
        return int(self.lib.nhw_stream(self.h) or 0)

    def debug_stop_after(self, label, occurrence=1):
        self._check(self.lib.nhw_debug_stop_after(self.h, label.encode() if label else None, occurrence), "debug_stop")

    def debug_read(self, what, img, dtype, count):
        a = np.zeros(count, dtype=dtype)
        self._check(self.lib.nhw_debug_read(self.h, what.encode(), img, a.ctypes.data, a.nbytes), "debug_read")
        return a

    def color_check(self):
        """mismatches between the integer and the IEEE colour transform over all 2^24 triples"""
        return int(self.lib.nhw_debug_color_check(self.h))

    def dec_color_check(self):
        """mismatches between the integer and the IEEE form of the decoder's q >= 20 colour matrix over all 2^24 triples"""
        return int(self.lib.nhw_debug_dec_color_check(self.h))

    def profile(self, mode):
        """0 off, 1 on, 2 on + reset"""
        self._check(self.lib.nhw_profile(self.h, int(mode)), "nhw_profile")

    def profile_table(self):
        """-> {kernel label: (total ms, launches)} accumulated since the last reset"""
        buf = ctypes.create_string_buffer(1 << 16)
        self.lib.nhw_profile_read(self.h, buf, len(buf))
        table = {}
        for line in buf.value.decode().splitlines():
            name, ms, cnt = line.split("\t")
            table[name] = (float(ms), int(cnt))
        return table

    # ---------------- host-buffer API (the reference-facing call) ----------------
    def encode(self, rgb, quality=20):
        """rgb: uint8 array (n, 786432) of raw BMP pixel bytes.  -> (list of bytes, status array)"""
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, PIX_BYTES)
        n = rgb.shape[0]
        out = np.empty(n * MAX_STREAM_BYTES, dtype=np.uint8)
        offs = np.zeros(n + 1, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        rc = self.lib.nhw_encode_batch(self.h, rgb.ctypes.data, n, int(quality), out.ctypes.data, out.size,
                                       offs.ctypes.data, status.ctypes.data)
        self._check(rc, "nhw_encode_batch")
        return [out[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(n)], status

    def encode_into(self, rgb, quality, out, offsets, status):
        """zero-allocation variant for bench.py: all arguments are preallocated numpy arrays"""
        n = rgb.shape[0]
        rc = self.lib.nhw_encode_batch(self.h, rgb.ctypes.data, n, int(quality), out.ctypes.data, out.size,
                                       offsets.ctypes.data, status.ctypes.data)
        self._check(rc, "nhw_encode_batch")

    def decode(self, streams):
        """streams: list of .nhw byte strings -> (uint8 array (n, 786432), status)"""
        n = len(streams)
        offs = np.zeros(n + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(s) for s in streams])
        blob = np.frombuffer(b"".join(streams), dtype=np.uint8)
        rgb = np.empty((n, PIX_BYTES), dtype=np.uint8)
        status = np.zeros(n, dtype=np.int32)
        rc = self.lib.nhw_decode_batch(self.h, blob.ctypes.data, offs.ctypes.data, n, rgb.ctypes.data,
                                       status.ctypes.data)
        self._check(rc, "nhw_decode_batch")
        return rgb, status

    def decode_into(self, blob, offsets, n, rgb, status):
        """zero-allocation variant for bench.py: blob/offsets as produced by encode_into"""
        rc = self.lib.nhw_decode_batch(self.h, blob.ctypes.data, offsets.ctypes.data, int(n), rgb.ctypes.data,
                                       status.ctypes.data)
        self._check(rc, "nhw_decode_batch")

    # ---------------- device-resident API (torch tensors on this GPU) ----------------
    def encode_device(self, rgb_t, quality, out_t, len_t, status_t):
        n = rgb_t.shape[0]
        rc = self.lib.nhw_encode_batch_device(self.h, _ptr(rgb_t), n, int(quality), _ptr(out_t), _ptr(len_t),
                                              _ptr(status_t))
        self._check(rc, "nhw_encode_batch_device")

    def pack_device(self, slots_t, len_t, offs_t, dense_t):
        """stream slots -> streams back to back (offs_t: n + 1 int64/uint64 on the device)"""
        rc = self.lib.nhw_pack_batch_device(self.h, _ptr(slots_t), _ptr(len_t), int(len_t.numel()), _ptr(offs_t), _ptr(dense_t))
        self._check(rc, "nhw_pack_batch_device")

    def digest_device(self, data_t, out_t, len_t=None):
        """one 64-bit checksum per row of data_t (2-D uint8, contiguous); len_t: optional per-row lengths"""
        n, stride = data_t.shape[0], data_t.shape[1]
        rc = self.lib.nhw_digest_batch_device(self.h, _ptr(data_t), int(stride), _ptr(len_t), int(stride), n, _ptr(out_t))
        self._check(rc, "nhw_digest_batch_device")

    def decode_device(self, in_t, len_t, rgb_t, status_t, stride=MAX_STREAM_BYTES):
        """streams in device memory, one per `stride`-byte slot (what encode_device writes) -> pixels in device memory"""
        n = rgb_t.shape[0]
        rc = self.lib.nhw_decode_batch_device(self.h, _ptr(in_t), int(stride), _ptr(len_t), n, _ptr(rgb_t), _ptr(status_t))
        self._check(rc, "nhw_decode_batch_device")

    def decode_packed_device(self, in_t, offs_t, rgb_t, status_t):
        """streams packed back to back in device memory (offs_t: n + 1 uint64/int64 offsets; 64 readable bytes must follow)"""
        n = rgb_t.shape[0]
        rc = self.lib.nhw_decode_batch_packed_device(self.h, _ptr(in_t), _ptr(offs_t), n, _ptr(rgb_t), _ptr(status_t))
        self._check(rc, "nhw_decode_batch_packed_device")

    def stage_frontend(self, rgb_t, quality, y_proc=None, y_ll1=None, c_proc=None, c_ll1=None):
        n = rgb_t.shape[0]
        rc = self.lib.nhw_stage_frontend_device(self.h, _ptr(rgb_t), n, int(quality), _ptr(y_proc), _ptr(y_ll1),
                                                _ptr(c_proc), _ptr(c_ll1))
        self._check(rc, "nhw_stage_frontend_device")

    def stage_colorspace(self, rgb_t, quality, pre, y=None, u=None, v=None):
        n = rgb_t.shape[0]
        rc = self.lib.nhw_stage_colorspace_device(self.h, _ptr(rgb_t), n, int(quality), int(bool(pre)), _ptr(y),
                                                  _ptr(u), _ptr(v))
        self._check(rc, "nhw_stage_colorspace_device")

    def synth(self, rgb_t, seed0, kind=0):
        import torch
        from . import synth as _synth
        if self._lut is None:
            self._lut = torch.from_numpy(_synth.sin_lut()).to(rgb_t.device)
        n = rgb_t.shape[0]
        rc = self.lib.nhw_synth_batch_device(self.h, _ptr(rgb_t), n, int(seed0) & 0xFFFFFFFF, int(kind),
                                             _ptr(self._lut))
        self._check(rc, "nhw_synth_batch_device")
