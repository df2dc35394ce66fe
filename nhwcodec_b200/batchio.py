"""ctypes binding of libnhw_batchio.so (include/nhw_batchio.h): image readers, 512x512 tiling, the .nhwpack container and
the manifest / directory batch jobs.  Host-side only; the codec work goes through a `Codec` context."""
import ctypes
import os

import numpy as np

from .capi import PIX_BYTES, load_library

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnhw_batchio.so")
EXPORTS = [
    "nhw_image_load", "nhw_image_load_mem", "nhw_image_free", "nhw_image_save", "nhw_tiles_x", "nhw_tiles_y",
    "nhw_image_to_tiles", "nhw_tiles_to_image", "nhw_pack_open", "nhw_pack_close", "nhw_pack_images", "nhw_pack_tiles",
    "nhw_pack_quality", "nhw_pack_image_info", "nhw_pack_image_name", "nhw_pack_tile_bytes", "nhw_pack_read_tiles",
    "nhw_batch_encode_files", "nhw_batch_encode_manifest", "nhw_batch_encode_dir", "nhw_batch_decode_pack",
    "nhw_batch_extract_pack", "nhw_batchio_last_error",
]


class BatchIOError(RuntimeError):
    pass


class _Image(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("pixels", ctypes.POINTER(ctypes.c_uint8))]


class PackImage(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("tiles_x", ctypes.c_uint32), ("tiles_y", ctypes.c_uint32),
                ("first_tile", ctypes.c_uint64), ("name_off", ctypes.c_uint32), ("name_len", ctypes.c_uint32)]


class Stats(ctypes.Structure):
    _fields_ = [("images", ctypes.c_uint64), ("tiles", ctypes.c_uint64), ("bytes_in", ctypes.c_uint64), ("bytes_out", ctypes.c_uint64),
                ("seconds_total", ctypes.c_double), ("seconds_read", ctypes.c_double), ("seconds_codec", ctypes.c_double),
                ("seconds_write", ctypes.c_double), ("first_bad_image", ctypes.c_int64), ("first_bad_status", ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        load_library()                     # libnhw_cuda.so first (the batch library links against it)
        if not os.path.exists(LIB_PATH):
            raise BatchIOError("libnhw_batchio.so is not built (run `python -m nhwcodec_b200.build`)")
        L = ctypes.CDLL(LIB_PATH)
        vp, u64, u32, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
        L.nhw_image_load.argtypes = [ctypes.c_char_p, ctypes.POINTER(_Image)]
        L.nhw_image_load_mem.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(_Image)]
        L.nhw_image_free.argtypes = [ctypes.POINTER(_Image)]
        L.nhw_image_free.restype = None
        L.nhw_image_save.argtypes = [ctypes.c_char_p, ctypes.POINTER(_Image), i32]
        L.nhw_tiles_x.argtypes = L.nhw_tiles_y.argtypes = [u32]
        L.nhw_tiles_x.restype = L.nhw_tiles_y.restype = u32
        L.nhw_image_to_tiles.argtypes = [ctypes.POINTER(_Image), vp]
        L.nhw_image_to_tiles.restype = None
        L.nhw_tiles_to_image.argtypes = [vp, ctypes.POINTER(_Image)]
        L.nhw_tiles_to_image.restype = None
        L.nhw_pack_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
        L.nhw_pack_close.argtypes = [vp]
        L.nhw_pack_close.restype = None
        L.nhw_pack_images.argtypes = L.nhw_pack_tiles.argtypes = [vp]
        L.nhw_pack_images.restype = L.nhw_pack_tiles.restype = u64
        L.nhw_pack_quality.argtypes = [vp]
        L.nhw_pack_image_info.argtypes = [vp, u64]
        L.nhw_pack_image_info.restype = ctypes.POINTER(PackImage)
        L.nhw_pack_image_name.argtypes = [vp, u64, ctypes.c_char_p, ctypes.c_size_t]
        L.nhw_pack_image_name.restype = ctypes.c_size_t
        L.nhw_pack_tile_bytes.argtypes = [vp, u64]
        L.nhw_pack_tile_bytes.restype = u64
        L.nhw_pack_read_tiles.argtypes = [vp, u64, u64, vp, ctypes.c_size_t, vp]
        L.nhw_batch_encode_files.argtypes = [vp, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p), u64, i32,
                                             ctypes.c_char_p, u32, ctypes.POINTER(Stats)]
        L.nhw_batch_encode_manifest.argtypes = [vp, ctypes.c_char_p, i32, ctypes.c_char_p, u32, ctypes.POINTER(Stats)]
        L.nhw_batch_encode_dir.argtypes = [vp, ctypes.c_char_p, i32, ctypes.c_char_p, u32, ctypes.POINTER(Stats)]
        L.nhw_batch_decode_pack.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, i32, u32, ctypes.POINTER(Stats)]
        L.nhw_batch_extract_pack.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(Stats)]
        L.nhw_batchio_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise BatchIOError("%s failed (%d): %s" % (what, rc, lib().nhw_batchio_last_error().decode()))


def load_image(path_or_bytes):
    """-> uint8 array (height, width, 3): B,G,R, bottom-up (row 0 = bottom line), the layout the codec consumes"""
    L, im = lib(), _Image()
    if isinstance(path_or_bytes, (bytes, bytearray)):
        buf = np.frombuffer(bytes(path_or_bytes), dtype=np.uint8)
        _check(L.nhw_image_load_mem(buf.ctypes.data, buf.size, ctypes.byref(im)), "nhw_image_load_mem")
    else:
        _check(L.nhw_image_load(os.fsencode(path_or_bytes), ctypes.byref(im)), "nhw_image_load")
    a = np.ctypeslib.as_array(im.pixels, shape=(im.height, im.width, 3)).copy()
    L.nhw_image_free(ctypes.byref(im))
    return a


def save_image(path, pixels, fmt=0):
    pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
    im = _Image(pixels.shape[1], pixels.shape[0], pixels.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    _check(lib().nhw_image_save(os.fsencode(path), ctypes.byref(im), int(fmt)), "nhw_image_save")


def to_tiles(pixels):
    """(h, w, 3) bottom-up BGR -> (tiles_y * tiles_x, 786432) tiles, row-major from the top-left"""
    L = lib()
    pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
    h, w = pixels.shape[:2]
    n = L.nhw_tiles_x(w) * L.nhw_tiles_y(h)
    out = np.empty((n, PIX_BYTES), dtype=np.uint8)
    im = _Image(w, h, pixels.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    L.nhw_image_to_tiles(ctypes.byref(im), out.ctypes.data)
    return out


def from_tiles(tiles, width, height):
    tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
    out = np.empty((height, width, 3), dtype=np.uint8)
    im = _Image(width, height, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    lib().nhw_tiles_to_image(tiles.ctypes.data, ctypes.byref(im))
    return out


class Pack:
    """an open .nhwpack (reading)"""

    def __init__(self, path):
        self.h = ctypes.c_void_p()
        _check(lib().nhw_pack_open(os.fsencode(path), ctypes.byref(self.h)), "nhw_pack_open")

    def close(self):
        if self.h:
            lib().nhw_pack_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def n_images(self):
        return int(lib().nhw_pack_images(self.h))

    @property
    def n_tiles(self):
        return int(lib().nhw_pack_tiles(self.h))

    @property
    def quality(self):
        return int(lib().nhw_pack_quality(self.h))

    def image(self, i):
        rec = lib().nhw_pack_image_info(self.h, i).contents
        buf = ctypes.create_string_buffer(4096)
        lib().nhw_pack_image_name(self.h, i, buf, len(buf))
        return {"width": rec.width, "height": rec.height, "tiles_x": rec.tiles_x, "tiles_y": rec.tiles_y,
                "first_tile": int(rec.first_tile), "name": buf.value.decode()}

    def tile(self, t):
        n = int(lib().nhw_pack_tile_bytes(self.h, t))
        buf = np.empty(max(n, 1), dtype=np.uint8)
        offs = np.zeros(2, dtype=np.uint64)
        _check(lib().nhw_pack_read_tiles(self.h, t, 1, buf.ctypes.data, buf.size, offs.ctypes.data), "nhw_pack_read_tiles")
        return buf[:n].tobytes()


def _paths(paths):
    arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    return arr


def encode_files(codec, paths, quality, pack_path, group_tiles=0, names=None):
    st = Stats()
    nm = _paths(names) if names else None
    _check(lib().nhw_batch_encode_files(codec.h, _paths(paths), nm, len(paths), int(quality), os.fsencode(pack_path), group_tiles,
                                        ctypes.byref(st)), "nhw_batch_encode_files")
    return st.as_dict()


def encode_manifest(codec, manifest, quality, pack_path, group_tiles=0):
    st = Stats()
    _check(lib().nhw_batch_encode_manifest(codec.h, os.fsencode(manifest), int(quality), os.fsencode(pack_path), group_tiles,
                                           ctypes.byref(st)), "nhw_batch_encode_manifest")
    return st.as_dict()


def encode_dir(codec, directory, quality, pack_path, group_tiles=0):
    st = Stats()
    _check(lib().nhw_batch_encode_dir(codec.h, os.fsencode(directory), int(quality), os.fsencode(pack_path), group_tiles,
                                      ctypes.byref(st)), "nhw_batch_encode_dir")
    return st.as_dict()


def decode_pack(codec, pack_path, out_dir, fmt=0, group_tiles=0):
    st = Stats()
    _check(lib().nhw_batch_decode_pack(codec.h, os.fsencode(pack_path), os.fsencode(out_dir), int(fmt), group_tiles, ctypes.byref(st)),
           "nhw_batch_decode_pack")
    return st.as_dict()


def extract_pack(pack_path, out_dir):
    st = Stats()
    _check(lib().nhw_batch_extract_pack(os.fsencode(pack_path), os.fsencode(out_dir), ctypes.byref(st)), "nhw_batch_extract_pack")
    return st.as_dict()
