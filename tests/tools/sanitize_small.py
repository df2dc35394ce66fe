"""a small encode + decode through both APIs, meant to be run under compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tests/tools/sanitize_small.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from nhwcodec_b200 import Codec, synth
c = Codec(device=0, max_batch=4)
imgs = np.stack([synth.natural(1), synth.textured(2), synth.noise(3), synth.natural(4)])
for q in (20, 23, 9):
    s, st = c.encode(imgs, q)
    assert (st == 0).all()
    back, dst = c.decode(s)
    assert (dst == 0).all()
t = torch.from_numpy(imgs).cuda()
out = torch.zeros((4, 1 << 19), dtype=torch.uint8, device="cuda"); ln = torch.zeros(4, dtype=torch.int32, device="cuda"); st = torch.zeros(4, dtype=torch.int32, device="cuda")
c.encode_device(t, 20, out, ln, st)
back = torch.empty((4, 786432), dtype=torch.uint8, device="cuda")
c.decode_device(out, ln, back, st)
torch.cuda.synchronize()
# pinned buffers: the SM-copy paths of the host API
pin = torch.from_numpy(imgs).pin_memory()
o2 = torch.empty(4 << 19, dtype=torch.uint8).pin_memory(); offs = np.zeros(5, dtype=np.uint64); s2 = np.zeros(4, dtype=np.int32)
c.encode_into(pin.numpy(), 20, o2.numpy(), offs, s2)
b2 = torch.empty((4, 786432), dtype=torch.uint8).pin_memory()
c.decode_into(o2.numpy(), offs, 4, b2.numpy(), s2)
assert (s2 == 0).all() and np.array_equal(b2.numpy(), back.cpu().numpy())
print("sanitize_small: done")
