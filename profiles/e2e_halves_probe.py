#!/usr/bin/env python
"""profiles/e2e_halves_probe.py -- does a second, staggered encode/decode pipeline over the other half of the batch keep the
link busier than one pipeline over the whole batch?  (A blocking call fills and drains its own sub-chunk pipeline: the
upload engine idles while the call's last kernels run, the download engine while its first ones do.)"""
import json
import os
import queue
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B, Q, STEPS, PIX = 4096, 20, 10, 262144


def main():
    import torch
    from nhwcodec_b200 import Codec
    from nhwcodec_b200.capi import PIX_BYTES
    torch.cuda.set_device(0)
    rgb_host = torch.empty((B, PIX_BYTES), dtype=torch.uint8, pin_memory=True)
    back_host = torch.empty((B, PIX_BYTES), dtype=torch.uint8, pin_memory=True)
    gen = Codec(device=0, max_batch=1024)
    rgb = torch.empty((B, PIX_BYTES), dtype=torch.uint8, device="cuda")
    gen.synth(rgb, 1000, 0)
    rgb_host.copy_(rgb)
    torch.cuda.synchronize()
    gen.close()
    del rgb
    rgb_np, back_np = rgb_host.numpy(), back_host.numpy()
    for parts, delay_ms in ((1, 0), (2, 0), (2, 25), (2, 40), (4, 12)):
        n = B // parts
        encs = [Codec(device=0, max_batch=n) for _ in range(parts)]
        decs = [Codec(device=0, max_batch=n) for _ in range(parts)]
        outs = [[torch.empty(n * 65536, dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)] for _ in range(parts)]
        offs = [[np.zeros(n + 1, dtype=np.uint64) for _ in range(2)] for _ in range(parts)]
        st, dst = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)

        def part(h, nsteps, delay):
            time.sleep(delay)
            lo, hi = h * n, (h + 1) * n
            q_free, q_ready = queue.Queue(), queue.Queue()
            q_free.put(0)
            q_free.put(1)

            def decoder():
                for _ in range(nsteps):
                    s = q_ready.get()
                    decs[h].decode_into(outs[h][s], offs[h][s], n, back_np[lo:hi], dst[lo:hi])
                    q_free.put(s)
            th = threading.Thread(target=decoder)
            th.start()
            for _ in range(nsteps):
                s = q_free.get()
                encs[h].encode_into(rgb_np[lo:hi], Q, outs[h][s], offs[h][s], st[lo:hi])
                q_ready.put(s)
            th.join()

        def run(nsteps):
            ths = [threading.Thread(target=part, args=(h, nsteps, h * delay_ms / 1e3)) for h in range(parts)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        run(2)
        t0 = time.perf_counter()
        run(STEPS)
        dt = (time.perf_counter() - t0) / STEPS
        assert (st == 0).all() and (dst == 0).all()
        print(json.dumps({"pipelines": parts, "images_each": n, "stagger_ms": delay_ms, "ms_per_step": round(dt * 1e3, 2),
                          "MPix_s": round(B * PIX / dt / 1e6, 1)}), flush=True)
        for c in encs + decs:
            c.close()
        del outs


if __name__ == "__main__":
    main()
