"""one low-quality encode of a small batch (for ncu captures of the q <= 16 kernels)"""
import sys, torch
sys.path.insert(0, '.')
from nhwcodec_b200 import Codec
from nhwcodec_b200.capi import PIX_BYTES
B, q = int(sys.argv[1]), int(sys.argv[2])
c = Codec(device=0, max_batch=B)
rgb = torch.empty((B, PIX_BYTES), dtype=torch.uint8, device='cuda'); c.synth(rgb, 3000, 0)
slots = torch.empty((B, 1 << 19), dtype=torch.uint8, device='cuda'); lens = torch.zeros(B, dtype=torch.int32, device='cuda'); st = torch.zeros(B, dtype=torch.int32, device='cuda')
for _ in range(2):
    c.encode_device(rgb, q, slots, lens, st)
torch.cuda.synchronize()
