"""Multi-GPU plumbing for batch encode/decode.

Images are independent units (all reference state lives in per-call structs, encoder/codec.h:112-181),
so a batch is cut into contiguous blocks, one per rank, and every rank runs the single-GPU path on
its block.  The data path has NO collective.  The only exchange is output assembly: the per-image
stream lengths are all-gathered, turned into offsets, and the variable-length streams are sent
point-to-point into their final place in rank `dst`'s buffer (NCCL has no gatherv; grouped
send/recv is the NCCL idiom and the same code runs on gloo for the CPU tests).

Decoded pixels are never gathered (262144 images = 206 GB): they stay sharded and only a digest
is reduced.
"""
from __future__ import annotations

import hashlib
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def partition(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of rank `rank`; blocks differ by at most one image."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    base, extra = divmod(n_total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def offsets_from_lengths(lengths: torch.Tensor) -> torch.Tensor:
    """Exclusive prefix sum, n+1 entries (the `offsets` array of nhw_encode_batch)."""
    out = torch.zeros(lengths.numel() + 1, dtype=torch.int64, device=lengths.device)
    torch.cumsum(lengths.to(torch.int64), 0, out=out[1:])
    return out


def gather_streams(dense: torch.Tensor, lengths: torch.Tensor, n_total: int, dst: int = 0,
                   group: Optional[dist.ProcessGroup] = None):
    """Assemble the streams of all ranks on rank `dst`.

    dense   : uint8, this rank's streams back to back (sum(lengths) bytes are used)
    lengths : this rank's per-image stream lengths (its block of `partition(n_total, ...)`)
    returns : (all_bytes uint8, offsets int64[n_total+1]) on `dst`, (None, offsets) elsewhere.
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    start, stop = partition(n_total, world, rank)
    if lengths.numel() != stop - start:
        raise ValueError("rank %d holds %d streams, its block has %d" % (rank, lengths.numel(), stop - start))
    # lengths: pad blocks to the largest block size so one all_gather serves uneven partitions
    per = -(-n_total // world)
    mine = torch.zeros(per, dtype=torch.int64, device=lengths.device)
    mine[: lengths.numel()] = lengths.to(torch.int64)
    allv = torch.empty(world * per, dtype=torch.int64, device=lengths.device)
    dist.all_gather_into_tensor(allv, mine, group=group)
    blocks = [allv[r * per: r * per + (partition(n_total, world, r)[1] - partition(n_total, world, r)[0])] for r in range(world)]
    all_len = torch.cat(blocks)
    offsets = offsets_from_lengths(all_len)
    block_bytes = [int(b.sum().item()) for b in blocks]
    block_off = np.concatenate([[0], np.cumsum(block_bytes)]).astype(np.int64)
    out = None
    ops = []
    if rank == dst:
        out = torch.empty(int(block_off[-1]), dtype=torch.uint8, device=dense.device)
        out[int(block_off[rank]): int(block_off[rank + 1])] = dense[: block_bytes[rank]]
        for r in range(world):
            if r != dst and block_bytes[r]:
                ops.append(dist.P2POp(dist.irecv, out[int(block_off[r]): int(block_off[r + 1])],
                                      dist.get_global_rank(group, r) if group is not None else r, group))
    elif block_bytes[rank]:
        ops.append(dist.P2POp(dist.isend, dense[: block_bytes[rank]],
                              dist.get_global_rank(group, dst) if group is not None else dst, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out, offsets


def digest_reduce(pixels: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None) -> str:
    """Digest of the sharded decoded pixels that does not depend on the number of ranks: every image
    (one row of `pixels`) is hashed on its own rank, the 16-byte digests are all-gathered in image
    order and hashed again (a checksum of checksums).  Equal between a 1-GPU and an N-GPU run."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    start, stop = partition(n_total, world, rank)
    host = pixels.detach().cpu().numpy().reshape(stop - start, -1)
    per = -(-n_total // world)
    mine = torch.zeros((per, 16), dtype=torch.uint8)
    for i in range(stop - start):
        mine[i] = torch.frombuffer(bytearray(hashlib.md5(host[i].tobytes()).digest()), dtype=torch.uint8)
    mine = mine.to(pixels.device)
    allv = torch.empty((world * per, 16), dtype=torch.uint8, device=pixels.device)
    dist.all_gather_into_tensor(allv, mine, group=group)
    allv = allv.cpu().numpy()
    h = hashlib.sha256()
    for r in range(world):
        a, b = partition(n_total, world, r)
        h.update(allv[r * per: r * per + (b - a)].tobytes())
    return h.hexdigest()
