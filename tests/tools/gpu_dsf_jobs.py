"""timing experiment: the decoder's serial front with one job at a time (results are wrong, only the kernel time is read)"""
import os, sys, json, torch
sys.path.insert(0, '.')
from nhwcodec_b200 import Codec
from nhwcodec_b200.capi import PIX_BYTES
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
enc = Codec(device=0, max_batch=B)
rgb = torch.empty((B, PIX_BYTES), dtype=torch.uint8, device='cuda'); enc.synth(rgb, 1000, 0)
slots = torch.empty((B, 1 << 19), dtype=torch.uint8, device='cuda'); lens = torch.zeros(B, dtype=torch.int32, device='cuda'); st = torch.zeros(B, dtype=torch.int32, device='cuda')
enc.encode_device(rgb, 20, slots, lens, st); torch.cuda.synchronize(); enc.close()
back = torch.empty((B, PIX_BYTES), dtype=torch.uint8, device='cuda')
for mask, spw in [(15, 4), (1, 4), (2, 4), (4, 4), (8, 4), (15, 2), (15, 1), (15, 8), (1, 1), (1, 2)]:
    os.environ['NHW_DSF_JOBS'] = str(mask); os.environ['NHW_DSF_STREAMS'] = str(spw)
    c = Codec(device=0, max_batch=B)
    c.decode_device(slots, lens, back, st)
    c.profile(2)
    for _ in range(3): c.decode_device(slots, lens, back, st)
    t = c.profile_table(); c.profile(0)
    print(json.dumps({"batch": B, "jobs": mask, "streams_per_warp": spw, "d_serial_front_ms": round(t['d_serial_front'][0] / t['d_serial_front'][1], 3)}), flush=True)
    c.close()
