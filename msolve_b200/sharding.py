"""Prime sharding of the QQ multi-modular loop across ranks (one rank per GPU).

msolve runs the application phase of its tracer for a batch of independent
primes in parallel (reference src/msolve/msolve.c:1815-1824, one prime per
OpenMP thread) and only needs the per-prime results on the host for CRT lifting
(msolve.c:2694-2740).  Across GPUs the same loop shards by prime: no exchange on
the data path, one gather of the results at the end of a batch.  The functions
here are backend agnostic (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def shard(items: Sequence[int], rank: int, world: int) -> List[int]:
    """Round-robin shard: item k goes to rank k % world (msolve hands prime i to
    thread i; with G GPUs prime i goes to GPU i % G)."""
    return list(items[rank::world])


def unshard(per_rank: Sequence[Sequence[int]]) -> List[int]:
    """Inverse of :func:`shard`: interleave the per-rank lists back into batch order."""
    world = len(per_rank)
    n = sum(len(x) for x in per_rank)
    out = [0] * n
    for r, xs in enumerate(per_rank):
        for k, v in enumerate(xs):
            out[r + k * world] = v
    return out


def gather_uint64(local: Sequence[int], dist=None, device="cpu") -> List[List[int]]:
    """All-gather one 64-bit value per local item (variable counts per rank).
    Returns the per-rank lists on every rank.  ``dist`` is torch.distributed (or
    None / uninitialised for a single process)."""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [list(int(x) for x in local)]
    import torch
    world = dist.get_world_size()
    n = torch.tensor([len(local)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    nmax = int(max(int(c.item()) for c in counts))
    # uint64 -> two non-negative int64 halves (NCCL / gloo have no uint64)
    arr = np.zeros((max(nmax, 1), 2), dtype=np.int64)
    for k, v in enumerate(local):
        arr[k, 0] = int(v) & 0xffffffff
        arr[k, 1] = (int(v) >> 32) & 0xffffffff
    t = torch.from_numpy(arr).to(device)
    bufs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(bufs, t)
    out = []
    for r in range(world):
        b = bufs[r].cpu().numpy()
        out.append([int(b[k, 0]) | (int(b[k, 1]) << 32) for k in range(int(counts[r].item()))])
    return out


def gather_residues(local: "np.ndarray", dist=None, device="cpu", out: "np.ndarray" = None, root: int = None) -> "np.ndarray":
    """All-gather the modular results of a batch: ``local`` is uint32 [k, words], one row per prime of this
    rank (the coefficients of the modular Groebner basis / parametrisation, what msolve hands to CRT,
    msolve.c:2694-2712).  Returns uint32 [sum k, words] in BATCH order (prime i of the batch came from rank
    i % world, see :func:`shard`) on every rank.  One collective for the counts, one for the payload; the
    interleave into batch order happens on the device, so the host sees one copy in and one copy out (at
    PCIe speed when ``local`` and ``out`` are page-locked, see :func:`pinned_rows`).  With ``root`` set only that
    rank copies the gathered block to its host (msolve lifts by CRT in one place); the others return None."""
    on_device = not isinstance(local, np.ndarray)       # a torch int32 tensor already in HBM (same bits as uint32)
    if not on_device:
        local = np.ascontiguousarray(local, dtype=np.uint32)
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        if not on_device:
            return local                                # nothing to exchange
        import torch
        if out is None:
            out = np.empty(tuple(local.shape), dtype=np.uint32)
        torch.from_numpy(out.view(np.int32)).copy_(local)
        return out
    import torch
    world = dist.get_world_size()
    k, words = tuple(local.shape)
    meta = torch.tensor([k, words], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    ks = [int(m[0].item()) for m in metas]
    assert all(int(m[1].item()) == words or int(m[0].item()) == 0 for m in metas), "ranks disagree on the residue length"
    # round-robin sharding: the counts differ by at most one and the longer shards come first
    assert all(ks[r] >= ks[r + 1] and ks[0] - ks[r + 1] <= 1 for r in range(world - 1)), "not a round-robin shard"
    kmax, total = max(max(ks), 1), sum(ks)
    words = max(words, 1)
    if on_device and k == kmax and tuple(local.shape) == (kmax, words) and local.is_contiguous():
        t = local                                                             # full shard already in HBM: no staging copy
    else:
        t = torch.zeros((kmax, words), dtype=torch.int32, device=device)      # NCCL / gloo: no uint32, same bits
        if k:
            t[:k].copy_(local if on_device else torch.from_numpy(local.view(np.int32)), non_blocking=True)
    big = torch.empty((world, kmax, words), dtype=torch.int32, device=device)
    dist.all_gather(list(big.unbind(0)), t)
    if root is not None and dist.get_rank() != root:
        return None
    inter = big.transpose(0, 1).reshape(world * kmax, words)[:total]          # row r + j * world = row j of rank r
    if out is None:
        out = np.empty((total, words), dtype=np.uint32)
    assert out.shape == (total, words) and out.dtype == np.uint32 and out.flags.c_contiguous
    torch.from_numpy(out.view(np.int32)).copy_(inter)
    return out


def pinned_rows(rows: int, words: int) -> "np.ndarray":
    """uint32 [rows, words] in page-locked host memory (the residues of a batch: written by the replays, read by the
    copy engine).  The torch tensor that owns the memory is kept alive by the array."""
    import torch
    t = torch.empty((max(rows, 1), max(words, 1)), dtype=torch.int32)
    if torch.cuda.is_available():
        t = t.pin_memory()
    a = t.numpy().view(np.uint32)[:rows, :words]
    return a
