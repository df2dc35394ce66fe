"""GPU-vs-host-emulation bisection: stop the CUDA encode after a labelled kernel and compare a
workspace plane with the host harness tap taken at the same pipeline point.

    python tests/hostemu/dbg_gpu.py <natural|noise|textured> <seed> <quality>      (needs a GPU)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

from nhwcodec_b200 import Codec, synth  # noqa: E402
from oracle import refbind  # noqa: E402
import run as R  # noqa: E402

kind, seed, q = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
pix = {"natural": synth.natural, "noise": synth.noise, "textured": synth.textured}[kind](seed)
_, rt = refbind.ref_encode_taps(pix, q)
stream, ht, order = R.host_encode(rt["y_pre"].view(np.int16), rt["cs_U"], rt["cs_V"], q)
c = Codec(0, 4)
# (label, occurrence, gpu array, host tap, region)
ONLY = os.environ.get("DBG_ONLY")
CHECKS = [
    ("k_dwt_level_smem<256>", 1, "proc", "y_dwt2_proc", None),
    ("y_e6a_tag", 1, "ll1", "y_e6a_ll1", None),
    ("y_recons1_quant", 1, "jpeg", "y_rec1_jpeg", 256),
    ("k_idwt_cols_t<256>", 1, "proc", "y_syn1_proc", None),
    ("y_e6c_apply", 1, "proc", "y_e6c_proc", None),
    ("y_e6d_correct", 1, "jpeg", "y_e6d_jpeg", 256),
    ("k_dwt_level_smem<256>", 2, "proc", "y_dwt2b_proc", None),
    ("y_recons0_shrink", 1, "jpeg", "y_rec0_jpeg", 256),
    ("k_idwt_cols_t<256>", 2, "proc", "y_syn0_proc", None),
    ("y_e14_e15_tags", 1, "proc", "y_e15_proc", None),
    ("y_e16_residual", 1, "proc", "y_e16_proc", None),
    ("y_e16_residual", 1, "ll1", "y_e16_ll1", None),
    ("y_e16b_classify", 1, "proc", "y_e16b_proc", None),
    ("y_e16b_classify", 1, "ll1", "y_e16b_ll1", None),
    ("y_e18_lists", 1, "ll1", "y_e18_ll1", None),
    ("y_e20_cleanup", 3, "proc", "y_e20_proc", None),
    ("y_offset_quant", 1, "proc", "y_e21_proc", None),
    ("y_scan", 1, "scan", "y_e23_scan", None),
    ("y_peephole", 1, "scan", "y_e24_scan", None),
]
for label, occ, what, tap, region in CHECKS:
    if ONLY and ONLY not in label:
        continue
    c.debug_stop_after(label, occ)
    try:
        c.encode(pix[None, :], q)
    except Exception:
        pass
    if what == "scan":
        g = c.debug_read("scan", 0, np.uint8, 262144)
        h = ht[tap]
        w = 2048
    else:
        n = 262144 if what in ("proc", "jpeg") else 65536
        g = c.debug_read(what, 0, np.int16, n)
        h = ht[tap].view(np.int16)
        w = 512 if n == 262144 else 256
        if region:
            g = g.reshape(512, 512)[:region, :region].reshape(-1)
            h = h.reshape(512, 512)[:region, :region].reshape(-1)
            w = region
    d = np.flatnonzero(g != h)
    print("%-24s #%d %-5s vs %-14s %s" % (label, occ, what, tap, "ok" if d.size == 0 else
          "DIFF %d cells, first (r=%d,c=%d) gpu %d host %d" % (d.size, d[0] // w, d[0] % w, g[d[0]], h[d[0]])))
    if d.size and what == "ll1":
        cols = np.unique(d % w)
        rows = np.unique(d // w)
        print("      cols:", cols[:40], "... rows:", rows[:10], "..", rows[-3:], "n=", d.size)
        print("      sample gpu/host:", [(int(g[i]), int(h[i])) for i in d[:12]])
c.debug_stop_after(None)
