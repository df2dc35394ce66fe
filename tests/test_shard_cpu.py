"""N>1 host logic on CPU: two gloo ranks shard a batch, assemble the streams on rank 0 and reduce a
digest.  The streams here come from the oracle (compiled reference) because there is no GPU in this
suite; the GPU path is the same code with the `nccl` backend (bench.py --gpus N)."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nhwcodec_b200 import shard


def test_partition_covers_everything():
    for n in (0, 1, 7, 8, 4096, 262144, 1001):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.partition(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for a, b in zip(blocks, blocks[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.partition(4, 2, 2)


def test_offsets_from_lengths():
    o = shard.offsets_from_lengths(torch.tensor([3, 0, 5], dtype=torch.int32))
    assert o.tolist() == [0, 3, 3, 8]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _streams(n):
    # deterministic variable-length "streams" (length 0 included: a failed image yields no bytes)
    rng = np.random.default_rng(7)
    lens = rng.integers(0, 5000, size=n)
    lens[3 % n] = 0
    return [rng.integers(0, 256, size=int(l), dtype=np.uint8) for l in lens]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        streams = _streams(n_total)
        a, b = shard.partition(n_total, world, rank)
        mine = streams[a:b]
        dense = torch.from_numpy(np.concatenate(mine + [np.zeros(16, np.uint8)]))   # slack after the used bytes
        lens = torch.tensor([len(s) for s in mine], dtype=torch.int32)
        out, offs = shard.gather_streams(dense, lens, n_total)
        pixels = torch.arange(a * 12, b * 12, dtype=torch.int64).reshape(b - a, 12).to(torch.uint8)
        dig = shard.digest_reduce(pixels, n_total)
        if rank == 0:
            q.put((out.numpy().tobytes(), offs.tolist(), dig))
        else:
            assert out is None
            q.put((None, offs.tolist(), dig))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 9), (2, 8), (3, 10)])
def test_gather_streams_gloo(world, n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    streams = _streams(n_total)
    want = np.concatenate(streams).tobytes()
    want_offs = np.concatenate([[0], np.cumsum([len(s) for s in streams])]).tolist()
    blobs = [g[0] for g in got if g[0] is not None]
    assert len(blobs) == 1 and blobs[0] == want
    for g in got:
        assert g[1] == want_offs
    # the digest equals the one a single rank computes over the whole batch
    pixels = torch.arange(0, n_total * 12, dtype=torch.int64).reshape(n_total, 12).to(torch.uint8).numpy()
    h = hashlib.sha256()
    for i in range(n_total):
        h.update(hashlib.md5(pixels[i].tobytes()).digest())
    assert all(g[2] == h.hexdigest() for g in got)
