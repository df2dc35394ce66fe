import torch, time, json
torch.cuda.init()
N = 1 << 30
res = {}
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
d_in = torch.empty(N, dtype=torch.uint8, device='cuda')
d_out = torch.empty(N, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=4):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    return reps * N / dt / 1e9
for _ in range(2):
    res['h2d_alone'] = run(True, False)
    res['d2h_alone'] = run(False, True)
    res['duplex_each'] = run(True, True)
# chunked duplex: 64 MB pieces
def run_chunked(reps=4, chunk=64<<20):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        for o in range(0, N, chunk):
            with torch.cuda.stream(s1): d_in[o:o+chunk].copy_(h_in[o:o+chunk], non_blocking=True)
            with torch.cuda.stream(s2): h_out[o:o+chunk].copy_(d_out[o:o+chunk], non_blocking=True)
    torch.cuda.synchronize()
    return reps * N / (time.perf_counter() - t) / 1e9
res['duplex_each_chunk64M'] = run_chunked()
print(json.dumps(res))
