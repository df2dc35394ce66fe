import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from nhwcodec_b200 import Codec, synth
from oracle import refbind
c = Codec(device=0, max_batch=16)
fs = [synth.natural, synth.textured, synth.noise]
for q in (1, 2, 5, 9):
    imgs = np.stack([fs[i % 3](9100 + q + i) for i in range(6)])
    streams = [refbind.ref_encode(imgs[i], q) for i in range(6)]
    bad = 0
    for rep in range(40):
        rgb, st = c.decode(streams)
        if (st != 0).any():
            bad += 1
            if bad <= 3: print("q", q, "rep", rep, "status", st.tolist(), flush=True)
        # interleave other work to vary the leftovers in the context
        if rep % 5 == 0:
            c.encode(imgs[:3], 20)
    print("q", q, "bad reps", bad, "of 40", flush=True)
