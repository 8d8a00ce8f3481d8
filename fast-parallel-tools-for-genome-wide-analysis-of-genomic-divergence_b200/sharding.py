"""Sharding a scan over the GPUs of one node: one process per GPU, contiguous genome ranges, no exchange
until the results are gathered once at the end.

The reference has a single shared-memory work queue (64 pthreads pulling tasks of 100 windows, each task also
reading the (wsize - wstep) halo to its right: statistics/fisher/threadfisher.c:106-116, 191-238). Windows are
independent, so across GPUs the same idea becomes: rank r owns the contiguous window range
[begin_r, end_r) and needs exactly the SNPs with position in [begin_r*wstep, (end_r-1)*wstep + wsize].
Random streams are keyed by the GLOBAL window index, so the gathered result is bit-identical for any
number of ranks (tests/test_sharding.py, tests/test_gpu_parity.py).
"""
import numpy as np

from . import api


def partition_windows(nwin, world):
    """contiguous, near-equal window ranges: [(begin, end)] * world"""
    base, extra = divmod(int(nwin), int(world))
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def snp_slice(pos, begin, end, wsize, wstep):
    """index range [lo, hi) of the SNPs a rank owning windows [begin, end) has to hold (its range plus the halo)"""
    if end <= begin:
        return 0, 0
    lo = int(np.searchsorted(pos, begin * wstep, side="left"))
    hi = int(np.searchsorted(pos, (end - 1) * wstep + wsize, side="right"))
    return lo, hi


def _gather(local_arrays, ranges, rank, world, group=None):
    """concatenate per-rank results on every rank; torch.distributed (nccl on GPUs, gloo in the CPU tests)"""
    if world == 1:
        return local_arrays
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    longest = max(e - b for b, e in ranges)
    out = []
    for a in local_arrays:
        pad = np.zeros(longest, dtype=a.dtype)
        pad[:a.size] = a
        mine = torch.from_numpy(pad).to(dev)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        out.append(np.concatenate([p.cpu().numpy()[:e - b] for p, (b, e) in zip(parts, ranges)]))
    return out


def fet_scan_sharded(a, b, pos, asize, bsize, regend, wsize, wstep, perc, rank, world, semantics=api.FPT_SCAN_SERIAL,
                     seed=None, compute=None, group=None):
    """FET scan of one chromosome split over `world` ranks; every rank returns the full (scores, stddev).
    `compute` defaults to the CUDA path (api.fet_scan); the CPU-only tests inject a checker with the same signature."""
    compute = compute or api.fet_scan
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    nwin = api.window_count(regend, wsize, wstep)
    ranges = partition_windows(nwin, world)
    wb, we = ranges[rank]
    lo, hi = snp_slice(pos, wb, we, wsize, wstep)
    s, d, _ = compute(a[lo * asize:hi * asize], b[lo * bsize:hi * bsize], pos[lo:hi], asize, bsize, regend, wsize, wstep, perc,
                      semantics=semantics, window_begin=wb, window_end=we, seed=seed)
    return tuple(_gather([s, d], ranges, rank, world, group))


def css_scan_sharded(a, b, pos, asize, bsize, regend, wsize, wstep, treshold, runs, rank, world, drosophila=0, mds=0,
                     semantics=api.FPT_SCAN_SERIAL, seed=None, compute=None, group=None):
    """CSS scan of one chromosome split over `world` ranks; every rank returns the full (scores, p)."""
    compute = compute or api.css_scan
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    nwin = api.window_count(regend, wsize, wstep)
    ranges = partition_windows(nwin, world)
    wb, we = ranges[rank]
    lo, hi = snp_slice(pos, wb, we, wsize, wstep)
    s, p, _ = compute(a[lo * asize:hi * asize], b[lo * bsize:hi * bsize], pos[lo:hi], asize, bsize, regend, wsize, wstep,
                      treshold, runs, drosophila=drosophila, mds=mds, semantics=semantics, window_begin=wb, window_end=we,
                      seed=seed)
    return tuple(_gather([s, p], ranges, rank, world, group))
