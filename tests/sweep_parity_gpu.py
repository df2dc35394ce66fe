"""tests/sweep_parity_gpu.py -- one-off wide parity sweep on a GPU box (not collected by pytest):
    python tests/sweep_parity_gpu.py [images per (kind, quality)] [lowest quality, default 1]
encodes kinds x qualities x seeds through libnhw_cuda and compares every stream with the compiled reference
(oracle/_ref, 16 host threads), then decodes a subset and compares the pixels with the reference decoder."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nhwcodec_b200 import Codec, synth  # noqa: E402
from oracle import refbind  # noqa: E402

per = int(sys.argv[1]) if len(sys.argv) > 1 else 32
codec = Codec(device=0, max_batch=128)
pool = ThreadPoolExecutor(16)
bad = 0
total = 0
q_lo = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for q in range(q_lo, 24):
    for kind, gen in (("natural", synth.natural), ("textured", synth.textured), ("noise", synth.noise)):
        seeds = [50000 + 997 * q + 13 * i for i in range(per)]
        imgs = np.stack(list(pool.map(gen, seeds)))
        streams, status = codec.encode(imgs, q)
        assert (status == 0).all(), (q, kind, status)
        want = list(pool.map(lambda im: refbind.ref_encode(im, q), imgs))
        miss = [seeds[i] for i in range(per) if streams[i] != want[i]]
        back, dstat = codec.decode(streams[:8])
        assert (dstat == 0).all()
        dmiss = [seeds[i] for i in range(8) if not np.array_equal(back[i], refbind.ref_decode(streams[i]))]
        total += per
        bad += len(miss) + len(dmiss)
        print("q%d %-8s %d streams, %d differ %s; 8 decodes, %d differ %s" % (q, kind, per, len(miss), miss[:4], len(dmiss), dmiss[:4]), flush=True)
print("SWEEP %s: %d encodes, %d mismatches" % ("OK" if bad == 0 else "FAILED", total, bad))
sys.exit(1 if bad else 0)
