"""Does an encode depend on what the context did before?  Slot 0 of a 1-image context replays the history of slot 3840 in the
1-GPU run of configs[4]: encode + decode of images 4000 + 3840 + 4096 k, k = 0..35; every stream is compared with the oracle."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from nhwcodec_b200 import Codec, synth, container
from oracle import refbind
K = int(sys.argv[1]) if len(sys.argv) > 1 else 36
seeds = [4000 + 3840 + 4096 * k for k in range(K)]
imgs = [synth.natural(s) for s in seeds]
want = [refbind.ref_encode(im, 20) for im in imgs]
c = Codec(device=0, max_batch=1)
bad = []
for k in range(K):
    s, st = c.encode(imgs[k][None, :], 20)
    if s[0] != want[k]:
        bad.append((k, container.first_difference(s[0], want[k])))
    c.decode([want[k]])
print("with decodes in between: bad", bad, flush=True)
c.close()
c = Codec(device=0, max_batch=1)
bad = []
for k in range(K):
    s, st = c.encode(imgs[k][None, :], 20)
    if s[0] != want[k]:
        bad.append((k, container.first_difference(s[0], want[k])))
print("encodes only: bad", bad, flush=True)
c.close()
# fresh context per image
bad = []
for k in range(K):
    c = Codec(device=0, max_batch=1)
    s, st = c.encode(imgs[k][None, :], 20)
    if s[0] != want[k]:
        bad.append((k, container.first_difference(s[0], want[k])))
    c.close()
print("fresh context each: bad", bad, flush=True)
