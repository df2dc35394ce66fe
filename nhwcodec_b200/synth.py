"""Host (numpy) mirror of csrc/synth.cu: the same bytes, integer arithmetic only.

natural(seed) / noise(seed) return the 786432 raw BMP pixel bytes of one synthetic image
(sinusoids + rectangles + grain, counter-based RNG).
"""
import numpy as np

M32 = 0xFFFFFFFF


def sin_lut():
    k = np.arange(1024)
    return np.round(256.0 * np.sin(2.0 * np.pi * k / 1024.0)).astype(np.int16)


def _fmix32(h):
    h = np.asarray(h, dtype=np.uint64) & M32
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M32
    h ^= h >> 16
    return h


def _prm(seed, k):
    return int(_fmix32((seed * 0x9E3779B1 + k * 0x632BE5AB + 0x7F4A7C15) & M32))


def _grain_hash(seed):
    pix = np.arange(512 * 512, dtype=np.uint64)[:, None]
    ch = np.arange(3, dtype=np.uint64)[None, :]
    a = (seed * 0x9E3779B1) & M32
    b = (((pix * 3 + ch) & M32) * 0x85EBCA77 + 0x1B873593) & M32
    return _fmix32(a ^ b)


def noise(seed):
    seed &= M32
    return (_grain_hash(seed) & 255).astype(np.uint8).reshape(-1)


def natural(seed, grain=1):
    seed &= M32
    lut = sin_lut().astype(np.int64)
    g = _grain_hash(seed).astype(np.int64)
    pix = np.arange(512 * 512, dtype=np.int64)
    x = (pix & 511)[:, None]
    y = (pix >> 9)[:, None]
    acc = np.full((512 * 512, 3), 128 << 8, dtype=np.int64)
    for t in range(12):
        p0, p1, p2, p3 = (_prm(seed, 4 * t + i) for i in range(4))
        fx = p0 % 25 - 12
        fy = p1 % 25 - 12
        ph = p2 & 1023
        amp = 4 + p3 % 24
        gain = np.array([128 + ((p3 >> (8 + 8 * ch)) & 127) for ch in range(3)], dtype=np.int64)[None, :]
        phase = ((fx * x + fy * y) * 2 + ph) & 1023
        acc += (amp * gain * lut[phase]) >> 8
    for t in range(10):
        q0, q1, q2, q3, q4 = (_prm(seed, 100 + 5 * t + i) for i in range(5))
        x0 = q0 & 511
        y0 = q1 & 511
        x1 = x0 + 16 + q2 % 200
        y1 = y0 + 16 + q3 % 200
        d = np.array([((q4 >> (8 * ch)) & 127) - 64 for ch in range(3)], dtype=np.int64)[None, :]
        inside = (x >= x0) & (x < x1) & (y >= y0) & (y < y1)
        acc += np.where(inside, d << 8, 0)
    s = np.zeros_like(g)
    for b in range(8):
        s += (g >> (2 * b)) & 3
    acc += ((s - 12) * grain) << 8
    val = np.clip(acc >> 8, 0, 255)
    return val.astype(np.uint8).reshape(-1)


def textured(seed):
    return natural(seed, grain=4)


def batch(seed0, n, kind=0):
    f = {0: natural, 1: noise, 2: textured}[kind]
    return np.stack([f(seed0 + i) for i in range(n)])
