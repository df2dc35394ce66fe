#!/usr/bin/env python
"""bench.py -- headline benchmark of the NHW hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's own CPU code)

Metric (BASELINE.json: "MPix/s encode+decode, 512x512 batch"): pixels taken through one full encode AND one full decode
per second.  One "step" = encode of one batch of synthetic images to .nhw streams, then decode of those streams back to
pixels (configs[1]: batch 4096 per GPU, -q20, natural-like generator of SURVEY.md section 8d).
  value : whole-job MPix/s with pixels and streams resident in HBM (device API: nhw_encode_batch_device +
          nhw_decode_batch_device), timed with CUDA events on the codec's stream, max over ranks.
  e2e   : the same round trip through the host-buffer C-ABI calls (nhw_encode_batch, nhw_decode_batch): pinned host
          pixels in, .nhw bytes to the host, back in, pixels out; every host<->device copy inside the timed region
          (and, with more than one GPU, the gather of all streams onto rank 0).
  roofline     : the kernel with the largest share of the step, algorithmic bytes per launch / its CUDA-event
                 duration per launch measured here, vs MEASURED_PEAKS.json; `decode.roofline` = the decoder's back end.
  cpu_baseline : oracle/_ref (the reference compiled with gcc -O3) doing the same round trip on the host cores,
                 bounded sample (rank 0, N=1 only).
Images shard trivially: each rank works on its own batch, no data-path collective (weak scaling).
The other BASELINE.json configs (decode-only 16384, quality sweep, 262144-image sharded round trip) are bench_configs.py.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PIX = 512 * 512
PIX_BYTES = PIX * 3
METRIC = "roundtrip_throughput_q20_512x512"
UNIT = "MPix/s"

# Algorithmic (compulsory) bytes per image moved by ONE launch of each kernel label; a label
# that is launched twice per step (both reconstructions, both analysis levels) moves these bytes each time, and its
# GB/s is computed per launch.
# planes: Y int16 512x512 = 524288 B, LL1 int16 256x256 = 131072 B, chroma int16 256x256 = 131072 B.
ALG_BYTES = {
    # fused front end: pixels in; three level-1 luma bands + LL1 (`res256`) + 4:2:0 chroma bytes out
    "k_front_luma": 786432 + 393216 + 131072 + 131072,
    "k_dwt_level<256>": 131072 + 131072,                # one level-2 analysis: LL1 in, four level-2 bands out
    "k_dwt_level<256,u8>": 2 * (65536 + 98304 + 32768), # chroma level 1 of both planes: bytes in, 3 bands + LL out
    "k_dwt_level<128>": 2 * (32768 + 32768),            # chroma level 2 of both planes
    # one synthesis level, band plane in + samples out.  <256> runs four times per round trip: three times on one luma
    # plane per image (2 x encode, 1 x decode) and once on the two chroma planes (decode, level 1): 1.25 planes per launch
    "k_idwt_level<256>": 2 * 131072 * 5 // 4,
    "k_idwt_level<128>": 2 * 2 * 32768,               # chroma level 2, both planes
    "y_quant_scan": 524288 + 262144,                   # coefficient plane in, scan bytes out
    "c_quant_scan": 2 * 131072 + 131072,
    "y_e6d_correct": 2 * 131072 + 2 * 131072,          # trial reconstruction + LL1 in, both corrected out
    "y_offset_mult8": 2 * 524288 * 3 // 4,             # the three level-1 bands in and out
    "y_e14_e15_tags": 2 * 524288 * 3 // 4,
    "y_offset_patterns": 2 * 131072,
    "y_e20_cleanup": 2 * 524288 * 3 // 4,
    "y_peephole": 2 * 262144,
    "entropy_pack": 2 * 393216,                        # two passes over the byte stream (+ output, added at run time)
    "y_e16_residual": 2 * 131072 + 131072,
    "y_e16b_classify": 2 * 131072 + 131072,
    "y_e18_lists": 3 * 131072,
    "y_recons_patterns": 2 * 131072,                   # level-2 region in, tags + im_jpeg samples out (one of the two launches)
    "y_recons0_quant": 2 * 98304,                      # level-2 detail bands in, im_jpeg out
    "y_recons1_quant": 2 * 98304,
    "y_e6a_tag": 98304 + 131072,
    "y_recons0_shrink": 2 * 131072,
    "y_ll2_code": 32768 + 3 * 16384,
    "c_ll_quant": 2 * (2 * 8192 + 4096 + 512),         # two 64x64 LL bands: read, zeroed, 4096 bytes + bit plane out
    # decoder
    "d_backend": 786432 + 786432,                      # band plane + two chroma planes (int16) in, BMP pixel bytes out
    "d_inv_rows_t": 131072 + 393216 + 524288,          # LL1 reconstruction + the three level-1 bands in, half-synthesised plane out
    "d_serial_front": 2 * 262144 + 2 * 131072 + 24576, # (+ stream bytes, added at run time) coefficient planes + LL bytes out
    "d_descan_y": 2 * 524288,
    "d_descan_uv": 2 * 2 * 131072,
    "d_sharpen_uv": 2 * 2 * 131072,
    "d_edge_flags": 2 * 131072,
    "d_markers_y": 2 * 524288,
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """samples before this point (warm-up) are dropped unless the timed region turns out too short to hold two"""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        first = getattr(self, "first", 0)
        window = "timed region"
        if len(self.lines) - first < 2:
            first, window = 0, "warm-up + timed region (timed region shorter than two samples)"
        self.window = window
        for ln in self.lines[first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": self.window, "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """The reference's own CPU code (oracle/_ref: gcc -O3 build of /root/reference, in memory, all host threads) on the
    same metric: every step encodes AND decodes `per_step` images of the GPU arm's batch (same generator, same seeds)."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from nhwcodec_b200 import synth
    from oracle import refbind
    cores = host_cores()
    distinct = min(64, 16 * cores)
    gen = [synth.natural, synth.noise, synth.textured][args.kind]
    images = np.stack([gen(1000 + i) for i in range(distinct)])      # the first seeds of the GPU arm's batch
    per_step = 16 * cores      # ~0.5 s of CPU work per step: long enough that thread start-up does not matter
    L = refbind.enc_stock_lib()
    D = refbind.dec_lib()
    streams = [np.frombuffer(refbind.ref_encode(images[i], args.quality), dtype=np.uint8).copy() for i in range(distinct)]
    scratch = [(np.zeros(PIX_BYTES, np.uint8), np.zeros(PIX_BYTES, np.uint8)) for _ in range(cores)]
    phase_s = {"enc": 0.0, "dec": 0.0}

    def step():
        idx = [0]
        lock = threading.Lock()

        def work(t):
            out, planes = scratch[t]
            while True:
                with lock:
                    i = idx[0]
                    idx[0] += 1
                if i >= per_step:
                    return
                k = i % distinct
                L.nhwref_encode_discard(images[k].ctypes.data, args.quality)
                D.nhwref_decode(streams[k].ctypes.data, streams[k].size, out.ctypes.data, planes.ctypes.data, 1)

        ths = [threading.Thread(target=work, args=(t,)) for t in range(cores)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps * per_step * PIX / dt / 1e6
    sample = ("%d images/step: %d distinct %s seeds 1000+ (the first of the GPU arm's batch), cycled; per image in-memory "
              "downsample_YUV420 + encode_image, then decode_image + write_image_bmp's pixel path (oracle/_ref, gcc -O3), %d threads"
              % (per_step, distinct, GENERATORS[args.kind], cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/f64", "data": "synthetic",
        "config": {"workload": workload_name(args.quality, per_step, args.kind) + " (reference CPU arm, bounded sample)",
                   "quality": args.quality, "images_per_step": per_step},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


GENERATORS = ["natural-like", "uniform-noise", "textured"]


def workload_name(q, batch, kind):
    return "round trip (encode, then decode) of a batch of %d synthetic 512x512 RGB images at -q%d per GPU, %s generator" % (
        batch, q, GENERATORS[kind])


def cpu_roundtrip_rate(images, quality, seconds, threads):
    """MPix/s of oracle/_ref on `threads` host threads: per image encode + decode, cycling over `images`."""
    from oracle import refbind
    L = refbind.enc_stock_lib()
    D = refbind.dec_lib()
    n = images.shape[0]
    streams = [np.frombuffer(refbind.ref_encode(images[i], quality), dtype=np.uint8).copy() for i in range(n)]
    stop_at = time.perf_counter() + seconds
    counts = [0] * threads
    enc_s = [0.0] * threads

    def work(t):
        out, planes = np.zeros(PIX_BYTES, np.uint8), np.zeros(PIX_BYTES, np.uint8)
        i = t
        while time.perf_counter() < stop_at:
            k = i % n
            a = time.perf_counter()
            rc = L.nhwref_encode_discard(images[k].ctypes.data, int(quality))
            enc_s[t] += time.perf_counter() - a
            if rc != 0:
                raise RuntimeError("reference encoder failed: %d" % rc)
            D.nhwref_decode(streams[k].ctypes.data, streams[k].size, out.ctypes.data, planes.ctypes.data, 1)
            counts[t] += 1
            i += threads

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    done = sum(counts)
    return done * PIX / dt / 1e6, done, dt, sum(enc_s) / max(dt * threads, 1e-9)


class Timer:
    """CUDA events on the codec's stream around a callable repeated `steps` times"""
    def __init__(self, torch, stream):
        self.torch, self.stream = torch, stream

    def __call__(self, fn, steps):
        t = self.torch
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        t.cuda.synchronize()
        return e0.elapsed_time(e1)


def kernel_rows(table, steps, batch, peak, mean_stream):
    total = sum(v[0] for v in table.values())
    rows = []
    for name, (kms, cnt) in sorted(table.items(), key=lambda kv: -kv[1][0]):
        per_launch_ms = kms / max(cnt, 1)
        b_img = ALG_BYTES.get(name)
        if b_img is not None and name in ("entropy_pack", "d_serial_front"):
            b_img = b_img + int(mean_stream)
        # a label launched more than once per step (both reconstructions, both analysis levels) moves its bytes each time
        gbs = (b_img * batch / (per_launch_ms / 1e3) / 1e9) if (b_img and per_launch_ms > 0) else None
        rows.append({"kernel": name, "ms_per_step": round(kms / steps, 4), "share": round(kms / total, 4) if total else None,
                     "launches_per_step": cnt // max(steps, 1), "alg_bytes_per_image_per_launch": b_img,
                     "GBps": round(gbs, 2) if gbs else None, "frac": round(gbs / peak, 5) if gbs else None})
    return rows


def ncu_traffic():
    """DRAM bytes per image of the kernels captured with `ncu --set full` this round: profiles/ncu_traffic.json, written
    by profiles/ncu_traffic.py from the committed capture (never typed in)"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def run_ours(args):
    import torch
    from nhwcodec_b200 import Codec

    rank, local, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B, q, K = args.batch, args.quality, args.steps
    codec = Codec(device=local, max_batch=B)
    stream = torch.cuda.ExternalStream(codec.stream_ptr, device=torch.device("cuda", local))
    timer = Timer(torch, stream)
    rgb = torch.empty((B, PIX_BYTES), dtype=torch.uint8, device="cuda")
    codec.synth(rgb, 1000 + rank * B, args.kind)
    slots = torch.empty((B, 1 << 19), dtype=torch.uint8, device="cuda")
    lens = torch.zeros(B, dtype=torch.int32, device="cuda")
    status = torch.zeros(B, dtype=torch.int32, device="cuda")
    back = torch.empty((B, PIX_BYTES), dtype=torch.uint8, device="cuda")
    dstatus = torch.zeros(B, dtype=torch.int32, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def enc():
        codec.encode_device(rgb, q, slots, lens, status)

    def dec():
        codec.decode_device(slots, lens, back, dstatus)

    def round_trip():
        enc()
        dec()

    # ---------------- device-resident round trip (value) ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()       # already streaming when the timed region starts; only its samples from then on are used
    for _ in range(args.warmup):
        round_trip()
    assert int((status != 0).sum().item()) == 0, "encode reported per-image errors"
    assert int((dstatus != 0).sum().item()) == 0, "decode reported per-image errors"
    mean_stream = float(lens.float().mean().item())
    err = (back[:64].float() - rgb[:64].float())
    psnr = float((10 * torch.log10(255.0 ** 2 / (err * err).mean(dim=1).clamp_min(1e-9))).min().item())
    barrier()
    codec.profile(0)          # the timed regions run un-instrumented
    launches0 = codec.launches
    sampler.mark()
    ms = timer(round_trip, K)
    launches = codec.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    ms_enc = timer(enc, K)
    barrier()
    ms_dec = timer(dec, K)
    # per-kernel table: the same steps again with CUDA events around every launch, one stream (serialised), so that
    # a kernel's time is its own; `share` is of this serialised pass
    codec.profile(2)
    for _ in range(K):
        round_trip()
    torch.cuda.synchronize()
    table = codec.profile_table()
    codec.profile(0)
    ms_max, ms_enc_max, ms_dec_max = reduce_max(ms), reduce_max(ms_enc), reduce_max(ms_dec)
    pix_job = world * B * K * PIX
    value = pix_job / (ms_max / 1e3) / 1e6

    # ---------------- secondary input distributions (device-resident round trip, rank 0's view) ----------------
    secondary = {}
    if not args.no_secondary:
        for kind in (1, 2):
            if kind == args.kind:
                continue
            codec.synth(rgb, 1000 + rank * B, kind)
            round_trip()
            m2 = float(lens.float().mean().item())
            ok = int((status != 0).sum().item()) + int((dstatus != 0).sum().item()) == 0
            t2 = reduce_max(timer(round_trip, 2))
            secondary[GENERATORS[kind]] = {"value": round(world * B * 2 * PIX / (t2 / 1e3) / 1e6, 3), "unit": UNIT, "steps": 2,
                                           "ms_per_step": round(t2 / 2, 3), "mean_stream_bytes": round(m2, 1), "errors": 0 if ok else 1}
        codec.synth(rgb, 1000 + rank * B, args.kind)
        round_trip()
        torch.cuda.synchronize()

    # ---------------- end to end through the host-buffer C-ABI (e2e) ----------------
    # pinned host pixels -> nhw_encode_batch -> .nhw bytes on the host -> nhw_decode_batch -> pixels on the host.
    # With more than one GPU the step also assembles every rank's streams on rank 0 (pack on the device, lengths
    # all-gather + point-to-point over NVLink): the one exchange step of the sharded job.
    rgb_host = torch.empty((B, PIX_BYTES), dtype=torch.uint8, pin_memory=True)
    rgb_host.copy_(rgb)
    cap = int(B * max(mean_stream * 1.5, 65536))
    out_host = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    back_host = torch.empty((B, PIX_BYTES), dtype=torch.uint8, pin_memory=True)
    offs = np.zeros(B + 1, dtype=np.uint64)
    st, dst = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    rgb_np, out_np, back_np = rgb_host.numpy(), out_host.numpy(), back_host.numpy()
    e2e_steps = max(1, min(K, 3))
    gather = None
    dense = offs_dev = None
    if dist is not None:
        from nhwcodec_b200 import shard
        dense = torch.empty(int(B * max(mean_stream * 1.5, 65536)) + 64, dtype=torch.uint8, device="cuda")
        offs_dev = torch.zeros(B + 1, dtype=torch.int64, device="cuda")

    def gather_step():
        codec.pack_device(slots, lens, offs_dev, dense)          # slots -> dense (k_pack_streams)
        return shard.gather_streams(dense, lens, world * B)

    def e2e_step():
        codec.encode_into(rgb_np, q, out_np, offs, st)
        if dist is not None:
            gather_step()
        codec.decode_into(out_np, offs, B, back_np, dst)

    e2e_step()                                                   # warm-up (and NCCL connection set-up)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_serial_s = reduce_max(time.perf_counter() - t0)
    # The same steps as a two-stage pipeline: an encoder context on this thread and a decoder context on a second host thread
    # (contexts are independent), two stream buffers in flight.  Step k's decode (download-heavy) overlaps step k+1's encode
    # (upload-heavy), so both directions of the PCIe link -- and the GPU -- stay busy; every step still uploads its own pixels
    # and downloads its own decoded pixels, and step k's decoder reads the bytes step k's encoder produced.  This is how a
    # caller streams batches through the blocking calls of the C-ABI.
    import queue
    import threading
    dec_codec = Codec(device=local, max_batch=B)
    out_b = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    outs, offs_l = [out_np, out_b.numpy()], [offs, np.zeros(B + 1, dtype=np.uint64)]
    pipe_steps = max(e2e_steps, min(3 * K, 30))       # the pipeline's fill and drain (one encode + one decode alone) are inside the timed region

    def run_pipeline(nsteps):
        q_free, q_ready = queue.Queue(), queue.Queue()
        q_free.put(0)
        q_free.put(1)
        err = []

        def decoder():
            try:
                torch.cuda.set_device(local)
                for _ in range(nsteps):
                    s_ = q_ready.get()
                    dec_codec.decode_into(outs[s_], offs_l[s_], B, back_np, dst)
                    q_free.put(s_)
            except Exception as e:  # surface it on the main thread
                err.append(e)
                q_free.put(0)
                q_free.put(1)

        th = threading.Thread(target=decoder)
        th.start()
        for _ in range(nsteps):
            s_ = q_free.get()
            codec.encode_into(rgb_np, q, outs[s_], offs_l[s_], st)
            if dist is not None:
                gather_step()
            q_ready.put(s_)
        th.join()
        if err:
            raise err[0]

    run_pipeline(2)
    barrier()
    t0 = time.perf_counter()
    run_pipeline(pipe_steps)
    torch.cuda.synchronize()
    e2e_s = reduce_max(time.perf_counter() - t0)
    dec_codec.close()
    del out_b
    assert int((st != 0).sum()) == 0 and int((dst != 0).sum()) == 0
    d2h_streams = int(offs[B])
    e2e_value = world * B * pipe_steps * PIX / e2e_s / 1e6
    # the two halves alone, for the record
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        codec.encode_into(rgb_np, q, out_np, offs, st)
    torch.cuda.synchronize()
    e2e_enc_s = reduce_max(time.perf_counter() - t0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        codec.decode_into(out_np, offs, B, back_np, dst)
    torch.cuda.synchronize()
    e2e_dec_s = reduce_max(time.perf_counter() - t0)
    if dist is not None:
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # events on torch's current stream: the point-to-point operations are ordered against it (the pack kernel runs on the
        # codec's stream inside a call that returns after it has finished)
        g0.record()
        allb, alloffs = gather_step()
        g1.record()
        torch.cuda.synchronize()
        barrier()
        tg = reduce_max(g0.elapsed_time(g1))
        total_bytes = int(alloffs[-1].item())
        gather = {"what": "k_pack_streams on every rank + lengths all-gather + point-to-point stream assembly on rank 0 (NCCL over "
                          "NVLink); inside the e2e timed region", "ms": round(tg, 3), "bytes_total": total_bytes,
                  "GBps_into_rank0": round(total_bytes * (world - 1) / world / (tg / 1e3) / 1e9, 2)}
        del allb
    # the same pinned pixels copied to the device with nothing else running, and back: the PCIe floor of one e2e step
    ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    rgb.copy_(rgb_host, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    ev0.record()
    rgb.copy_(rgb_host, non_blocking=True)
    ev1.record()
    back_host.copy_(back, non_blocking=True)
    ev2.record()
    torch.cuda.synchronize()
    h2d_alone_ms, d2h_alone_ms = reduce_max(ev0.elapsed_time(ev1)), reduce_max(ev1.elapsed_time(ev2))

    if rank == 0:
        peak, peak_src = peaks()
        kernels = kernel_rows(table, K, B, peak, mean_stream)
        by = {k["kernel"]: k for k in kernels}
        traffic = ncu_traffic()

        def roof(k):
            if not k:
                return None
            tr = traffic.get(k["kernel"])
            return {"bound": "hbm", "kernel": k["kernel"], "achieved": k["GBps"], "peak": peak, "unit": "GB/s", "frac": k["frac"],
                    "traffic": int(tr["dram_bytes_per_image"] * B) if tr else None,
                    "traffic_source": tr["source"] if tr else None, "peak_source": peak_src,
                    "ms_per_launch": round(k["ms_per_step"] / max(k["launches_per_step"], 1), 4), "share_of_step": k["share"]}

        dom = next((k for k in kernels if k["GBps"]), None)          # largest share of the step among the kernels with a byte model
        enc_ms = sum(k["ms_per_step"] for k in kernels if not k["kernel"].startswith(("d_", "kd_")))
        dec_ms = sum(k["ms_per_step"] for k in kernels if k["kernel"].startswith(("d_", "kd_")))
        # the north-star work: k_front_luma + ONE of the k_dwt_level<256> launches (the other is the closed loop's) + the
        # chroma levels (k_dwt_level<128> likewise runs twice per step, once here)
        front_ms = sum(by[nm]["ms_per_step"] * part for nm, part in (("k_front_luma", 1.0), ("k_dwt_level<256>", 0.5),
                                                                      ("k_dwt_level<256,u8>", 1.0), ("k_dwt_level<128>", 0.5)) if nm in by)
        frontend = None
        if front_ms > 0:
            gbs = 1572864 * B / (front_ms / 1e3) / 1e9
            frontend = {"what": "the north-star 'fused colorspace+DWT' work: k_front_luma (colour + 4:2:0 + pre-sharpen + level-1 luma "
                                "DWT in one pass) + luma level 2 + chroma levels 1-2 (k_dwt_level); 6 B/pixel (SURVEY 8d)",
                        "ms_per_step": round(front_ms, 4), "alg_bytes_per_image": 1572864, "GBps": round(gbs, 2), "frac": round(gbs / peak, 5),
                        "k_front_luma": roof(by.get("k_front_luma"))}
        whole = {"encode": {"alg_bytes_per_image": PIX_BYTES + int(mean_stream), "ms_per_step": round(ms_enc_max / K, 4),
                            "GBps": round((PIX_BYTES + mean_stream) * B / (ms_enc_max / K / 1e3) / 1e9, 2)},
                 "decode": {"alg_bytes_per_image": PIX_BYTES + int(mean_stream), "ms_per_step": round(ms_dec_max / K, 4),
                            "GBps": round((PIX_BYTES + mean_stream) * B / (ms_dec_max / K / 1e3) / 1e9, 2)}}
        for w in whole.values():
            w["frac"] = round(w["GBps"] / peak, 5)

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cores = host_cores()
                sample_imgs = rgb[:min(B, 32)].cpu().numpy()
                v, done, secs, enc_share = cpu_roundtrip_rate(sample_imgs, q, args.cpu_seconds, cores)
                cpu = {"value": round(v, 3), "unit": UNIT, "cores": cores, "kind": "reference",
                       "sample": "%d round trips (encode + decode) of the first %d images of this batch in %.1f s, in memory "
                                 "(oracle/_ref, gcc -O3), %d threads; encode is %.0f %% of that time" % (
                                     done, sample_imgs.shape[0], secs, cores, 100 * enc_share)}
            except Exception as e:  # oracle/_ref missing on this box
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}

        e2e = {"value": round(e2e_value, 3), "unit": UNIT,
               "h2d_bytes_per_step": B * PIX_BYTES + d2h_streams + 8 * (B + 1),
               "d2h_bytes_per_step": d2h_streams + 8 * (B + 1) + 4 * B + B * PIX_BYTES + 4 * B, "steps": pipe_steps,
               "ms_per_step": round(e2e_s * 1e3 / pipe_steps, 3),
               "what": "per step: nhw_encode_batch (pinned host pixels -> .nhw bytes on the host), then nhw_decode_batch (those bytes -> "
                       "pixels on the host)%s; every copy inside the timed region.  The steps run as a two-stage pipeline (encoder context on "
                       "one host thread, decoder context on another, two stream buffers): step k's decode overlaps step k+1's encode, "
                       "so both PCIe directions are busy; wall time from the first upload to the last download / number of steps" % (
                           ", then the stream gather onto rank 0" if dist is not None else ""),
               "one_context_serial": {"value": round(world * B * e2e_steps * PIX / e2e_serial_s / 1e6, 3),
                                      "ms_per_step": round(e2e_serial_s * 1e3 / e2e_steps, 3),
                                      "steps": e2e_steps,
                                      "what": "the same step as two blocking calls on one context (encode all, then decode all), no overlap between steps"},
               "encode_only": {"value": round(world * B * e2e_steps * PIX / e2e_enc_s / 1e6, 3), "ms_per_step": round(e2e_enc_s * 1e3 / e2e_steps, 3)},
               "decode_only": {"value": round(world * B * e2e_steps * PIX / e2e_dec_s / 1e6, 3), "ms_per_step": round(e2e_dec_s * 1e3 / e2e_steps, 3)},
               "pcie_floor": {"h2d_pixels_ms": round(h2d_alone_ms, 3), "d2h_pixels_ms": round(d2h_alone_ms, 3),
                              "h2d_GBps": round(B * PIX_BYTES / h2d_alone_ms / 1e6, 2), "d2h_GBps": round(B * PIX_BYTES / d2h_alone_ms / 1e6, 2),
                              "note": "max over ranks of one copy of this rank's pixels alone, all ranks copying at once"},
               "gather": gather}
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": round(ms_max / K, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16 (integer colour; f64 on exact ties and in the decoder's colour matrix)",
            "data": "synthetic",
            "config": {"workload": workload_name(q, B, args.kind) + " (BASELINE.json configs[1] batch, encode + decode as the metric asks)",
                       "quality": q, "batch_per_gpu": B, "generator": GENERATORS[args.kind],
                       "mean_stream_bytes": round(mean_stream, 1), "min_psnr_db_first_64": round(psnr, 2),
                       "bit_exact": "verified by tests/test_encode_gpu.py, tests/test_decode_gpu.py (q1..23)",
                       "value_is": "pixels through one full encode + one full decode per second, pixels and streams resident in HBM",
                       "l2": "inputs (%.1f GB per step) exceed the 126 MB L2; no flush needed" % (B * PIX_BYTES / 1e9),
                       "encode_only": {"value": round(pix_job / (ms_enc_max / 1e3) / 1e6, 3), "ms_per_step": round(ms_enc_max / K, 4)},
                       "decode_only": {"value": round(pix_job / (ms_dec_max / 1e3) / 1e6, 3), "ms_per_step": round(ms_dec_max / K, 4)},
                       "secondary_inputs": secondary},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "kernel_table": "per-kernel ms from a separate serialised pass (CUDA events around every launch on the codec's stream); "
                            "in the e2e region the host-buffer calls run 16 (encode) / 8 (decode) sub-chunks on 4 streams",
            "clocks": clocks,
            "roofline": roof(dom),
            "frontend": frontend,
            "decode": {"value": round(pix_job / (ms_dec_max / 1e3) / 1e6, 3), "unit": UNIT, "ms_per_step": round(ms_dec_max / K, 4),
                       "kernel_ms_per_step": round(dec_ms, 3), "roofline": roof(by.get("d_backend")),
                       "what": "nhw_decode_batch_device on the %d streams of the encode above (headers walked on the device)" % B},
            "encode": {"value": round(pix_job / (ms_enc_max / 1e3) / 1e6, 3), "unit": UNIT, "ms_per_step": round(ms_enc_max / K, 4),
                       "kernel_ms_per_step": round(enc_ms, 3)},
            "whole_path_roofline": whole,
            "cpu_baseline": cpu,
            "kernels": kernels[:60],
        }
        print(json.dumps(line), flush=True)
    codec.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="images per GPU per step")
    ap.add_argument("--quality", type=int, default=20)
    ap.add_argument("--kind", type=int, default=0, help="0 natural-like, 1 uniform noise, 2 textured")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the uniform-noise / textured secondary measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
