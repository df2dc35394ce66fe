"""tests/hostemu/fuzz_decode.py -- TEST TOOLING: mutated .nhw streams through the host build of the decoder's stage
functions under AddressSanitizer.  The decoder must refuse or decode every input without touching memory outside
its workspace (the arrays here are sized like the device workspace: planes with their guard bands, 65536-entry lists).

    python tests/hostemu/fuzz_decode.py <seed> <mutations per stream>      (re-executes itself with libasan preloaded)
"""
import ctypes
import os
import random
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
ASAN_SO = os.path.join(HERE, "libhostemu_asan.so")


def build():
    src = os.path.join(HERE, "hostemu.cpp")
    if os.path.exists(ASAN_SO) and os.path.getmtime(ASAN_SO) > os.path.getmtime(src):
        csrc = os.path.join(ROOT, "nhwcodec_b200", "csrc")
        if all(os.path.getmtime(ASAN_SO) > os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))):
            return
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=address", "-Wno-unused-function",
                           "-Wno-sign-compare", "-ffp-contract=off", src, "-o", ASAN_SO])


def mutate(good, rnd):
    b = bytearray(good)
    mode = rnd.randrange(4)
    if mode == 0:      # a header field
        off = rnd.randrange(2, 50)
        struct.pack_into("<H", b, off, rnd.choice([0, 1, 255, 4096, 8192, 65535, rnd.randrange(65536)]))
        b += b"\0" * rnd.choice([0, 100000, 600000])
    elif mode == 1:    # scattered bytes
        for _ in range(rnd.randrange(1, 200)):
            b[rnd.randrange(len(b))] = rnd.randrange(256)
    elif mode == 2:    # truncation
        b = b[:rnd.randrange(len(b))]
    else:              # a stretch of extreme bytes
        a = rnd.randrange(50, len(b))
        e = min(len(b), a + rnd.randrange(1, 20000))
        for i in range(a, e):
            b[i] = rnd.choice([255, 128, 127, rnd.randrange(256)])
    return bytes(b)


def main():
    seed, count = int(sys.argv[1]), int(sys.argv[2])
    if os.environ.get("NHW_FUZZ_CHILD") != "1":
        build()
        asan = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
        env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0", NHW_FUZZ_CHILD="1")
        sys.exit(subprocess.call([sys.executable, __file__, str(seed), str(count)], env=env))
    from nhwcodec_b200 import synth
    from oracle import refbind
    L = ctypes.CDLL(ASAN_SO)
    L.he_decode.restype = ctypes.c_int
    L.he_decode.argtypes = [ctypes.c_char_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p]
    rgb = np.zeros(786432, dtype=np.uint8)
    yuv = np.zeros(786432, dtype=np.uint8)
    rnd = random.Random(seed)
    ok = total = 0
    for q in (23, 19, 12, 21, 16, 5):
        good = refbind.ref_encode(synth.textured(40 + q), q)
        for _ in range(count):
            s = mutate(good, rnd)
            rc = L.he_decode(s + b"\0" * 64, len(s), rgb.ctypes.data, yuv.ctypes.data)
            ok += rc == 0
            total += 1
    print("fuzz: %d streams, %d decoded, %d refused, no memory error" % (total, ok, total - ok))


if __name__ == "__main__":
    main()
