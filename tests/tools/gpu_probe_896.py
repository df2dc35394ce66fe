import sys, numpy as np, torch
sys.path.insert(0, '.')
from nhwcodec_b200 import Codec, synth, container
from oracle import refbind
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pre = len(sys.argv) > 2
if pre:   # what the other tests do first: small contexts, host API, decode
    c0 = Codec(device=0, max_batch=16)
    imgs = np.stack([synth.natural(5000 + i) for i in range(4)])
    s, st = c0.encode(imgs, 20); c0.decode(s); s, st = c0.encode(imgs, 9); c0.decode(s)
codec = Codec(device=0, max_batch=N)
rgb = torch.empty((N, 786432), dtype=torch.uint8, device='cuda')
codec.synth(rgb, 1000, 0)
slots = torch.zeros((N, 1 << 19), dtype=torch.uint8, device='cuda'); lens = torch.zeros(N, dtype=torch.int32, device='cuda'); st = torch.zeros(N, dtype=torch.int32, device='cuda')
for rep in range(3):
    codec.encode_device(rgb, 20, slots, lens, st)
    l = lens.cpu().numpy()
    bad = []
    for i in list(range(0, N, 128)) + [895, 897]:
        s = slots[i, :int(l[i])].cpu().numpy().tobytes()
        want = refbind.ref_encode(synth.natural(1000 + i), 20)
        if s != want: bad.append((i, container.first_difference(s, want)))
    print("rep", rep, "bad", bad, flush=True)
