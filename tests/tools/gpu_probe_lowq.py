import sys, numpy as np, torch
sys.path.insert(0,'.')
from nhwcodec_b200 import Codec, synth, container
from oracle import refbind
codec=Codec(device=0,max_batch=1024)
rgb=torch.empty((1024,786432),dtype=torch.uint8,device='cuda')
codec.synth(rgb,3000,0)
slots=torch.empty((1024,1<<19),dtype=torch.uint8,device='cuda'); lens=torch.zeros(1024,dtype=torch.int32,device='cuda'); st=torch.zeros(1024,dtype=torch.int32,device='cuda')
for q in (4,9,12):
    outs=[]
    for rep in range(3):
        codec.encode_device(rgb,q,slots,lens,st)
        l=lens.cpu().numpy()
        dig=torch.zeros(1024,dtype=torch.int64,device='cuda'); codec.digest_device(slots,dig,lens)
        outs.append(dig.cpu().numpy().copy())
    print(q,"deterministic:",[int((outs[0]!=o).sum()) for o in outs[1:]])
    bad=[]
    for i in list(range(500,540))+[0,1,2,3]:
        s=slots[i,:int(l[i])].cpu().numpy().tobytes()
        want=refbind.ref_encode(rgb[i].cpu().numpy(),q)
        if s!=want: bad.append((i,container.first_difference(s,want)))
    print(q,"bad",bad)
