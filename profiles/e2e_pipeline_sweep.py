#!/usr/bin/env python
"""profiles/e2e_pipeline_sweep.py -- the end-to-end round trip of bench.py (pinned host pixels -> nhw_encode_batch ->
nhw_decode_batch -> pinned host pixels, encoder and decoder contexts as a two-stage pipeline) for several settings of the
sub-chunk knobs, plus the two device-resident calls run side by side (no PCIe) to separate GPU sharing from link sharing.
One JSON line per setting."""
import json
import os
import queue
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B, Q, STEPS = 4096, 20, 10
PIX = 262144


def main():
    import torch
    from nhwcodec_b200 import Codec
    from nhwcodec_b200.capi import PIX_BYTES
    torch.cuda.set_device(0)
    rgb_host = torch.empty((B, PIX_BYTES), dtype=torch.uint8, pin_memory=True)
    back_host = torch.empty((B, PIX_BYTES), dtype=torch.uint8, pin_memory=True)
    cap = B * 65536
    outs = [torch.empty(cap, dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
    offs = [np.zeros(B + 1, dtype=np.uint64) for _ in range(2)]
    st, dst = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    settings = [dict(), dict(NHW_SUBS_ENCODE="8"), dict(NHW_SUBS_ENCODE="8", NHW_SUBS_DECODE="4"),
                dict(NHW_SUBS_ENCODE="8", NHW_LANES_ENCODE="2", NHW_SUBS_DECODE="4", NHW_LANES_DECODE="2"),
                dict(NHW_SUBS_ENCODE="4", NHW_LANES_ENCODE="2", NHW_SUBS_DECODE="4", NHW_LANES_DECODE="2"),
                dict(NHW_SUBS_ENCODE="16", NHW_SUBS_DECODE="16")]
    if len(sys.argv) > 1:
        settings = [json.loads(a) for a in sys.argv[1:]]
    for env in settings:
        for k, v in env.items():
            os.environ[k] = v
        enc, dec = Codec(device=0, max_batch=B), Codec(device=0, max_batch=B)
        for k in env:
            del os.environ[k]
        rgb = torch.empty((B, PIX_BYTES), dtype=torch.uint8, device="cuda")
        enc.synth(rgb, 1000, 0)
        rgb_host.copy_(rgb)
        torch.cuda.synchronize()
        rgb_np, back_np = rgb_host.numpy(), back_host.numpy()

        def pipeline(n):
            q_free, q_ready = queue.Queue(), queue.Queue()
            q_free.put(0)
            q_free.put(1)

            def decoder():
                for _ in range(n):
                    s = q_ready.get()
                    dec.decode_into(outs[s], offs[s], B, back_np, dst)
                    q_free.put(s)
            th = threading.Thread(target=decoder)
            th.start()
            for _ in range(n):
                s = q_free.get()
                enc.encode_into(rgb_np, Q, outs[s], offs[s], st)
                q_ready.put(s)
            th.join()
        pipeline(2)
        t0 = time.perf_counter()
        pipeline(STEPS)
        t_pipe = (time.perf_counter() - t0) / STEPS
        t0 = time.perf_counter()
        for _ in range(3):
            enc.encode_into(rgb_np, Q, outs[0], offs[0], st)
        t_enc = (time.perf_counter() - t0) / 3
        t0 = time.perf_counter()
        for _ in range(3):
            dec.decode_into(outs[0], offs[0], B, back_np, dst)
        t_dec = (time.perf_counter() - t0) / 3
        # device-resident: encode and decode side by side on the two contexts, nothing crosses the link
        slots = torch.empty((B, 1 << 19), dtype=torch.uint8, device="cuda")
        lens = torch.zeros(B, dtype=torch.int32, device="cuda")
        s1, s2 = torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda")
        back = torch.empty((B, PIX_BYTES), dtype=torch.uint8, device="cuda")
        slots2, lens2 = slots.clone(), lens.clone()
        enc.encode_device(rgb, Q, slots, lens, s1)
        slots2.copy_(slots)
        lens2.copy_(lens)
        torch.cuda.synchronize()

        def both(n):
            def d():
                for _ in range(n):
                    dec.decode_device(slots2, lens2, back, s2)
            th = threading.Thread(target=d)
            th.start()
            for _ in range(n):
                enc.encode_device(rgb, Q, slots, lens, s1)
            th.join()
            torch.cuda.synchronize()
        both(1)
        t0 = time.perf_counter()
        both(5)
        t_both = (time.perf_counter() - t0) / 5
        assert (st == 0).all() and (dst == 0).all()
        print(json.dumps({"env": env, "pipeline_ms": round(t_pipe * 1e3, 2), "MPix_s": round(B * PIX / t_pipe / 1e6, 1),
                          "encode_alone_ms": round(t_enc * 1e3, 2), "decode_alone_ms": round(t_dec * 1e3, 2),
                          "device_resident_side_by_side_ms": round(t_both * 1e3, 2)}), flush=True)
        enc.close()
        dec.close()
        del slots, slots2, back, rgb


if __name__ == "__main__":
    main()
