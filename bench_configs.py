#!/usr/bin/env python
"""bench_configs.py -- the BASELINE.json configurations other than the headline one (bench.py runs configs[1]).

    python bench_configs.py --config 2                      decode-only, batch 16384, 1 GPU
    python bench_configs.py --config 3                      quality sweep -q1..-q23, batch 1024, 1 GPU
    torchrun --nproc-per-node N bench_configs.py --config 4 [--total 262144]
                                                            round trip of `total` images sharded over N GPUs

Every configuration prints ONE JSON line on rank 0.  Inputs follow SURVEY.md section 8(d): natural-like generator,
seeds 2000+i (config 2), 3000+i (config 3), 4000+i (config 4), generated on the device.  Timing: CUDA events on the
codec's stream, inputs resident in HBM (the PCIe-inclusive number of the headline configuration is bench.py's e2e).
Parity inside these runs: config 3 compares two streams per quality with the compiled reference (oracle/_ref) when it
is present; config 4 compares a 1-in-64 subsample with it and prints a digest of ALL streams and ALL decoded pixels
that does not depend on the number of GPUs (equal digests at N=1 and N=8 = identical results).
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bench import ALG_BYTES, PIX, PIX_BYTES, UNIT, Timer, env_int, kernel_rows, peaks  # noqa: E402


def oracle():
    try:
        from oracle import refbind
        return refbind if refbind.available() else None
    except Exception:
        return None


def setup(batch):
    import torch
    from nhwcodec_b200 import Codec
    rank, local, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    codec = Codec(device=local, max_batch=batch)
    stream = torch.cuda.ExternalStream(codec.stream_ptr, device=torch.device("cuda", local))
    return torch, dist, codec, Timer(torch, stream), rank, local, world


def buffers(torch, B):
    return dict(rgb=torch.empty((B, PIX_BYTES), dtype=torch.uint8, device="cuda"),
                slots=torch.empty((B, 1 << 19), dtype=torch.uint8, device="cuda"),
                lens=torch.zeros(B, dtype=torch.int32, device="cuda"),
                st=torch.zeros(B, dtype=torch.int32, device="cuda"),
                back=torch.empty((B, PIX_BYTES), dtype=torch.uint8, device="cuda"),
                dst=torch.zeros(B, dtype=torch.int32, device="cuda"))


def config2(args):
    """decode-only batch 16384 .nhw -> BMP pixels on one GPU: the streams of our encoder (bit-exact with the canonical
    reference encoder, tests/test_encode_gpu.py) for seeds 2000..6095 at -q20, each decoded 4 times per step"""
    B = 4096
    torch, dist, codec, timer, rank, local, world = setup(B)
    b = buffers(torch, B)
    codec.synth(b["rgb"], 2000, 0)
    codec.encode_device(b["rgb"], args.quality, b["slots"], b["lens"], b["st"])
    assert int((b["st"] != 0).sum()) == 0
    mean_stream = float(b["lens"].float().mean())

    def step():
        for _ in range(4):
            codec.decode_device(b["slots"], b["lens"], b["back"], b["dst"])

    for _ in range(3):
        step()
    assert int((b["dst"] != 0).sum()) == 0
    ms = timer(step, args.steps)
    codec.profile(2)
    step()
    torch.cuda.synchronize()
    table = codec.profile_table()
    codec.profile(0)
    peak, peak_src = peaks()
    rows = kernel_rows(table, 4, B, peak, mean_stream)
    ref = oracle()
    checked = 0
    if ref is not None:
        lens = b["lens"].cpu().numpy()
        for i in range(0, B, 512):
            s = b["slots"][i, : int(lens[i])].cpu().numpy().tobytes()
            assert np.array_equal(b["back"][i].cpu().numpy(), ref.ref_decode(s)), i
            checked += 1
    value = 16384 * args.steps * PIX / (ms / 1e3) / 1e6
    be = next((r for r in rows if r["kernel"] == "d_backend"), None)
    print(json.dumps({
        "config": "BASELINE.json configs[2]: decode-only batch 16384 .nhw -> BMP pixels on 1xB200 (IDWT + filters + YCbCr->RGB)",
        "metric": "decode_throughput_q%d_512x512" % args.quality, "value": round(value, 3), "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "ms_per_step": round(ms / args.steps, 3), "streams_per_step": 16384,
        "distinct_streams": B, "mean_stream_bytes": round(mean_stream, 1), "oracle_checked_images": checked,
        "whole_decode": {"alg_bytes_per_image": PIX_BYTES + int(mean_stream),
                         "GBps": round((PIX_BYTES + mean_stream) * 16384 / (ms / args.steps / 1e3) / 1e9, 2),
                         "frac": round((PIX_BYTES + mean_stream) * 16384 / (ms / args.steps / 1e3) / 1e9 / peak, 5)},
        "roofline": None if not be else {"bound": "hbm", "kernel": "d_backend", "achieved": be["GBps"], "peak": peak, "unit": "GB/s",
                                         "frac": be["frac"], "peak_source": peak_src},
        "kernels": rows[:30]}), flush=True)
    codec.close()


def config3(args):
    """quality sweep -q1..-q23, batch 1024, seeds 3000+i: encode and decode time per quality, stream size, the three
    most expensive kernels of each direction"""
    B = 1024
    torch, dist, codec, timer, rank, local, world = setup(B)
    b = buffers(torch, B)
    codec.synth(b["rgb"], 3000, 0)
    ref = oracle()
    peak, _ = peaks()
    sweep = []
    for q in range(1, 24):
        def enc():
            codec.encode_device(b["rgb"], q, b["slots"], b["lens"], b["st"])

        def dec():
            codec.decode_device(b["slots"], b["lens"], b["back"], b["dst"])

        enc(); dec(); enc(); dec()
        assert int((b["st"] != 0).sum()) == 0 and int((b["dst"] != 0).sum()) == 0, q
        ms_e, ms_d = timer(enc, args.steps), timer(dec, args.steps)
        codec.profile(2)
        enc(); dec()
        torch.cuda.synchronize()
        table = codec.profile_table()
        codec.profile(0)
        mean_stream = float(b["lens"].float().mean())
        rows = kernel_rows(table, 1, B, peak, mean_stream)
        top_e = [(r["kernel"], r["ms_per_step"]) for r in rows if not r["kernel"].startswith(("d_", "kd_"))][:3]
        top_d = [(r["kernel"], r["ms_per_step"]) for r in rows if r["kernel"].startswith(("d_", "kd_"))][:3]
        err = b["back"][:32].float() - b["rgb"][:32].float()
        psnr = float((10 * torch.log10(255.0 ** 2 / (err * err).mean(dim=1).clamp_min(1e-9))).mean())
        checked = 0
        if ref is not None:
            lens = b["lens"].cpu().numpy()
            for i in (0, 517):
                s = b["slots"][i, : int(lens[i])].cpu().numpy().tobytes()
                assert s == ref.ref_encode(b["rgb"][i].cpu().numpy(), q), (q, i)
                assert np.array_equal(b["back"][i].cpu().numpy(), ref.ref_decode(s)), (q, i)
                checked += 1
        sweep.append({"q": q, "encode_ms": round(ms_e / args.steps, 3), "decode_ms": round(ms_d / args.steps, 3),
                      "encode_MPix_s": round(B * PIX / (ms_e / args.steps / 1e3) / 1e6, 1),
                      "decode_MPix_s": round(B * PIX / (ms_d / args.steps / 1e3) / 1e6, 1),
                      "mean_stream_bytes": round(mean_stream, 1), "mean_psnr_db": round(psnr, 2),
                      "encode_GBps": round((PIX_BYTES + mean_stream) * B / (ms_e / args.steps / 1e3) / 1e9, 1),
                      "top_encode_kernels": top_e, "top_decode_kernels": top_d, "oracle_checked_images": checked})
    print(json.dumps({"config": "BASELINE.json configs[3]: quality sweep -q1..-q23, batch 1024, 1xB200", "unit": UNIT, "batch": B,
                      "steps": args.steps, "sweep": sweep}), flush=True)
    codec.close()


def config4(args):
    """round trip of `total` images sharded over the GPUs of one box: every rank generates, encodes, decodes and digests
    its contiguous block in chunks; streams are gathered onto rank 0 chunk by chunk (the one exchange step, timed);
    decoded pixels stay sharded, only digests travel"""
    B = args.chunk
    torch, dist, codec, timer, rank, local, world = setup(B)
    from nhwcodec_b200 import shard
    total = args.total
    start, stop = shard.partition(total, world, rank)
    b = buffers(torch, B)
    dense = torch.empty(B * 131072 + 64, dtype=torch.uint8, device="cuda")
    offs_dev = torch.zeros(B + 1, dtype=torch.int64, device="cuda")
    dig_s = torch.zeros(stop - start, dtype=torch.int64, device="cuda")
    dig_p = torch.zeros(stop - start, dtype=torch.int64, device="cuda")
    ref = oracle() if not args.no_oracle else None
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=max(1, (os.cpu_count() or 8) // world))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_codec = t_gather = 0.0
    bytes_streams = 0
    checked = 0
    if dist is not None:
        # NCCL sets its point-to-point connections up on first use (hundreds of milliseconds for seven peers): once, here
        warm = torch.zeros(4096, dtype=torch.uint8, device="cuda")
        shard.gather_streams(warm, torch.full((4,), 1024, dtype=torch.int32, device="cuda"), 4 * world)
        dist.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for c0 in range(start, stop, B):
        n = min(B, stop - c0)
        v = {k: t[:n] for k, t in b.items()}
        codec.synth(v["rgb"], 4000 + c0, 0)
        ev[0].record(timer.stream)
        codec.encode_device(v["rgb"], args.quality, v["slots"], v["lens"], v["st"])
        codec.decode_device(v["slots"], v["lens"], v["back"], v["dst"])
        ev[1].record(timer.stream)
        codec.digest_device(v["slots"], dig_s[c0 - start: c0 - start + n], v["lens"])
        codec.digest_device(v["back"], dig_p[c0 - start: c0 - start + n])
        torch.cuda.synchronize()
        assert int((v["st"] != 0).sum()) == 0 and int((v["dst"] != 0).sum()) == 0, c0
        t_codec += ev[0].elapsed_time(ev[1])
        bytes_streams += int(v["lens"].sum())
        if dist is not None:   # all ranks walk their blocks in lockstep (equal block sizes), so the gathers line up
            ev[2].record()
            codec.pack_device(v["slots"], v["lens"], offs_dev[: n + 1], dense)
            gathered, _ = shard.gather_streams(dense, v["lens"], n * world)
            ev[3].record()
            torch.cuda.synchronize()
            t_gather += ev[2].elapsed_time(ev[3])
            del gathered
        if ref is not None:
            lens = v["lens"].cpu().numpy()
            idx = list(range((-c0) % 64, n, 64))     # global image index a multiple of 64
            pix = v["rgb"][idx].cpu().numpy()
            got = [v["slots"][i, : int(lens[i])].cpu().numpy().tobytes() for i in idx]
            dec = {i: v["back"][i].cpu().numpy() for i in idx if (c0 + i) % 1024 == 0}

            def verify(k):
                i = idx[k]
                want = ref.ref_encode(pix[k], args.quality)
                if got[k] != want:
                    return (c0 + i, "stream", want)
                if i in dec and not np.array_equal(dec[i], ref.ref_decode(got[k])):
                    return (c0 + i, "pixels", None)
                return None

            # (the compiled reference releases the GIL: host threads in parallel)
            for k, res in enumerate(pool.map(verify, range(len(idx)))):
                if res is not None:
                    # say what differs, and ask the oracle a second time from this thread: an answer that changes between two
                    # calls on the same pixels is the oracle's problem (the reference reads never-written memory), not the GPU's
                    from nhwcodec_b200 import container
                    again = ref.ref_encode(pix[k], args.quality)
                    where = container.first_difference(got[k], res[2]) if res[2] is not None else "decoded pixels"
                    raise AssertionError("image %d (%s): GPU vs oracle differ at %s; second oracle call %s the first, %s the GPU" % (
                        res[0], res[1], where, "equals" if again == res[2] else "DIFFERS from",
                        "equals" if again == got[k] else "differs from"))
            checked += len(idx)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall = time.perf_counter() - wall0

    def allcat(x):
        if dist is None:
            return x.cpu().numpy()
        per = -(-total // world)
        mine = torch.zeros(per, dtype=torch.int64, device="cuda")
        mine[: x.numel()] = x
        allv = torch.empty(world * per, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allv, mine)
        allv = allv.cpu().numpy()
        return np.concatenate([allv[r * per: r * per + (shard.partition(total, world, r)[1] - shard.partition(total, world, r)[0])]
                               for r in range(world)])

    all_s, all_p = allcat(dig_s), allcat(dig_p)
    red = torch.tensor([t_codec, t_gather, float(bytes_streams), float(checked)], dtype=torch.float64, device="cuda")
    mx = red.clone()
    if dist is not None:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(red, op=dist.ReduceOp.SUM)
    if rank == 0:
        codec_ms, gather_ms = float(mx[0]), float(mx[1])
        print(json.dumps({
            "config": "BASELINE.json configs[4]: batch %d encode+decode round trip sharded over %d x B200" % (total, world),
            "metric": "roundtrip_throughput_q%d_512x512" % args.quality, "unit": UNIT, "n_gpus": world, "images": total,
            "chunk": B, "value": round(total * PIX / (codec_ms / 1e3) / 1e6, 3),
            "value_is": "all images / max over ranks of the summed CUDA-event time of encode + decode (device-resident)",
            "codec_ms_max_rank": round(codec_ms, 1),
            "gather": None if dist is None else {"ms_max_rank": round(gather_ms, 1), "bytes_total": int(float(red[2])),
                                                 "GBps_into_rank0": round(float(red[2]) * (world - 1) / world / (gather_ms / 1e3) / 1e9, 2),
                                                 "what": "per chunk: k_pack_streams + lengths all-gather + point-to-point assembly on rank 0 (NCCL)"},
            "value_with_gather": round(total * PIX / ((codec_ms + gather_ms) / 1e3) / 1e6, 3),
            "wall_s_incl_generation_digests_oracle": round(wall, 2),
            "oracle_checked_images": int(float(red[3])), "oracle_rule": "every 64th image: stream bytes; every 1024th: decoded pixels too",
            "digest_streams_sha256": hashlib.sha256(all_s.tobytes()).hexdigest(),
            "digest_pixels_sha256": hashlib.sha256(all_p.tobytes()).hexdigest(),
            "digest_is": "sha256 over the per-image 64-bit checksums (nhw_digest_batch_device) in image order: independent of the sharding",
        }), flush=True)
    codec.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[2, 3, 4])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--quality", type=int, default=20)
    ap.add_argument("--total", type=int, default=262144)
    ap.add_argument("--chunk", type=int, default=4096)
    ap.add_argument("--no-oracle", action="store_true")
    args = ap.parse_args()
    {2: config2, 3: config3, 4: config4}[args.config](args)


if __name__ == "__main__":
    main()
