"""timing experiment: the q <= 16 pre-sharpening walker with one walk at a time (results are wrong, only the kernel time is read)"""
import os, sys, json, torch
sys.path.insert(0, '.')
from nhwcodec_b200 import Codec
from nhwcodec_b200.capi import PIX_BYTES
B = 1024
for q in (4, 9, 15):
    for mask in (15, 1, 2, 4, 8):
        os.environ['NHW_PLW_PHASES'] = str(mask)
        c = Codec(device=0, max_batch=B)
        rgb = torch.empty((B, PIX_BYTES), dtype=torch.uint8, device='cuda'); c.synth(rgb, 3000, 0)
        slots = torch.empty((B, 1 << 19), dtype=torch.uint8, device='cuda'); lens = torch.zeros(B, dtype=torch.int32, device='cuda'); st = torch.zeros(B, dtype=torch.int32, device='cuda')
        c.encode_device(rgb, q, slots, lens, st)
        c.profile(2)
        c.encode_device(rgb, q, slots, lens, st)
        t = c.profile_table(); c.profile(0)
        print(json.dumps({"q": q, "phases": mask, "k_pre_low_walk_ms": round(t['k_pre_low_walk'][0] / t['k_pre_low_walk'][1], 2)}), flush=True)
        c.close(); del rgb, slots
