"""Multi-rank host logic on the CPU: contiguous window ranges + halo SNPs per rank, world_size-2 gloo gather, and
the invariant that the gathered result does not depend on the number of ranks (random streams are keyed by the global
window index). The compute function is injected: here the oracle stands in for the CUDA path, which the GPU suite
exercises with the same sharding code."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_fet(a, b, pos, asize, bsize, regend, wsize, wstep, perc, semantics=0, window_begin=None, window_end=None, seed=None):
    import checkers
    from checkers import dptr, iptr
    o = checkers.load_oracle()
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    apos, bpos = np.repeat(pos, asize).astype(np.int32), np.repeat(pos, bsize).astype(np.int32)
    n = regend // wstep
    s, d = np.zeros(n), np.zeros(n)
    if pos.size:
        assert o.fpt_oracle_fet_scan(dptr(a), dptr(b), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, a.size, b.size, perc,
                                     dptr(s), dptr(d), semantics, seed) == 0
    return s[window_begin:window_end].copy(), d[window_begin:window_end].copy(), None


def _oracle_css(a, b, pos, asize, bsize, regend, wsize, wstep, treshold, runs, drosophila=0, mds=0, semantics=0, window_begin=None,
                window_end=None, seed=None):
    import checkers
    from checkers import dptr, iptr
    o = checkers.load_oracle()
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    apos, bpos = np.repeat(pos, asize).astype(np.int32), np.repeat(pos, bsize).astype(np.int32)
    n = regend // wstep
    s, p = np.zeros(n), np.zeros(n)
    if pos.size:
        assert o.fpt_oracle_css_scan(dptr(a), dptr(b), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, a.size, b.size, treshold, runs,
                                     drosophila, mds, dptr(s), dptr(p), semantics, seed) == 0
    return s[window_begin:window_end].copy(), p[window_begin:window_end].copy(), None


def _data():
    import fpt_b200.synth as synth
    ch = synth.chromosome(77, 80000, 1600, 6, 5)
    av, bv, _, _ = synth.reference_layout(ch)
    return ch, av, bv


def test_partition_and_halo():
    from fpt_b200.sharding import partition_windows, snp_slice
    for nwin, world in ((160, 1), (160, 2), (161, 4), (7, 8), (0, 3)):
        r = partition_windows(nwin, world)
        assert len(r) == world and r[0][0] == 0 and r[-1][1] == nwin
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    pos = np.array([0, 499, 500, 2999, 3000, 3001, 9000], dtype=np.int32)
    assert snp_slice(pos, 0, 1, 2500, 500) == (0, 3)              # window 0 = [0, 2500]
    assert snp_slice(pos, 1, 2, 2500, 500) == (2, 5)              # window 1 = [500, 3000], right edge inclusive
    assert snp_slice(pos, 3, 3, 2500, 500) == (0, 0)


def test_single_process_sharding_is_rank_count_invariant():
    from fpt_b200.sharding import css_scan_sharded, fet_scan_sharded, partition_windows
    ch, av, bv = _data()
    regend, wsize, wstep = 80000, 2500, 500
    full_s, full_d = fet_scan_sharded(av, bv, ch["pos"], 6, 5, regend, wsize, wstep, 0.95, 0, 1, seed=5, compute=_oracle_fet)
    for world in (2, 3):
        ranges = partition_windows(regend // wstep, world)
        s_cat, d_cat = [], []
        for rank in range(world):          # emulate the ranks one after the other (no process group: gather is the identity)
            from fpt_b200 import sharding
            wb, we = ranges[rank]
            lo, hi = sharding.snp_slice(ch["pos"], wb, we, wsize, wstep)
            s, d, _ = _oracle_fet(av[lo * 6:hi * 6], bv[lo * 5:hi * 5], ch["pos"][lo:hi], 6, 5, regend, wsize, wstep, 0.95,
                                  window_begin=wb, window_end=we, seed=5)
            s_cat.append(s)
            d_cat.append(d)
        assert np.array_equal(np.concatenate(s_cat), full_s) and np.array_equal(np.concatenate(d_cat), full_d)
    cs, cp = css_scan_sharded(av, bv, ch["pos"], 6, 5, regend, wsize, wstep, 5, 40, 0, 1, seed=5, compute=_oracle_css)
    assert (cp != 0).sum() > 50 and cs.size == regend // wstep


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from fpt_b200.sharding import css_scan_sharded, fet_scan_sharded
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        ch, av, bv = _data()
        s, d = fet_scan_sharded(av, bv, ch["pos"], 6, 5, 80000, 2500, 500, 0.95, rank, world, seed=5, compute=_oracle_fet)
        cs, cp = css_scan_sharded(av, bv, ch["pos"], 6, 5, 80000, 2500, 500, 5, 40, rank, world, seed=5, compute=_oracle_css)
        q.put((rank, s, d, cs, cp))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_gather_matches_single_rank():
    import torch.multiprocessing as mp
    from fpt_b200.sharding import css_scan_sharded, fet_scan_sharded
    ch, av, bv = _data()
    s1, d1 = fet_scan_sharded(av, bv, ch["pos"], 6, 5, 80000, 2500, 500, 0.95, 0, 1, seed=5, compute=_oracle_fet)
    c1, p1 = css_scan_sharded(av, bv, ch["pos"], 6, 5, 80000, 2500, 500, 5, 40, 0, 1, seed=5, compute=_oracle_css)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, s, d, cs, cp in got:            # every rank ends up with the full, identical result
        assert np.array_equal(s, s1) and np.array_equal(d, d1)
        assert np.array_equal(cs, c1, equal_nan=True) and np.array_equal(cp, p1)


def test_bench_genome_pieces_tile_the_genome():
    """bench.py --gpus N splits ONE genome into N contiguous window ranges (chromosomes dealt contiguously, the boundary
    chromosome cut): every window of every chromosome belongs to exactly one rank, in order, for every N the bench is run at"""
    import importlib
    bench = importlib.import_module("bench")
    nchrom, nout = 21, 42857
    for world in (1, 2, 3, 4, 8):
        seen = []
        prev_end = 0
        for rank in range(world):
            pieces, (gb, ge) = bench.genome_pieces(nchrom, nout, rank, world)
            assert gb == prev_end
            prev_end = ge
            assert sum(e - b for _, b, e in pieces) == ge - gb
            for c, b, e in pieces:
                assert 0 <= b < e <= nout
                seen.append((c * nout + b, c * nout + e))
        assert prev_end == nchrom * nout
        seen.sort()
        assert seen[0][0] == 0 and all(a[1] == b[0] for a, b in zip(seen, seen[1:])) and seen[-1][1] == nchrom * nout


def test_bench_large_cohort_work_models_are_per_chromosome():
    """bench.py's large-cohort kernel table: a chromosome of 2600 windows ran the code route in three passes of <= 1024 windows when
    that line was taken (one profile scope each; a chromosome is one pass since), so the Lanczos and GEMM work of the whole chromosome is set against the sum of the passes, not
    against one pass. Fed with the profile of the committed round-2 line (2 steps x 8 chromosomes on one GPU)."""
    import importlib
    bench = importlib.import_module("bench")
    prof = {"css_pack": {"ms": 7.248, "launches": 16}, "window_table": {"ms": 0.362, "launches": 16},
            "css_k4": {"ms": 76.33, "launches": 48}, "css_mds_large": {"ms": 1739.28, "launches": 48},
            "css_observed": {"ms": 110.82, "launches": 16}, "css_perm": {"ms": 841.88, "launches": 16}}
    large = {"_prof": prof, "_nwin": 2600, "_chrom_per_rank": 8, "_nsteps": 2, "_lanczos_steps": 73.6}
    micro = {"fp64": {"tflops": 33.783}, "umma_i8": {"tops": 4525.1}}
    out = bench.large_kernel_table(large, micro, 6552.6)
    assert not [k for k in out if k.startswith("_")]
    k = out["kernels"]
    assert abs(k["css_mds_large"]["ms_per_chromosome"] - 108.7) < 0.05 and k["css_mds_large"]["scopes_per_chromosome"] == 3
    assert abs(k["css_mds_large"]["ms_per_launch"] - 36.235) < 0.01
    fl = 2600 * 73.6 * (2.0e6 + 2000.0 * 73.6)
    assert abs(k["css_mds_large"]["fp64"]["algorithmic_flops"] - fl) < 1e-6 * fl
    assert abs(k["css_mds_large"]["fp64"]["frac"] - fl / 108.705e-3 / 33.783e12) < 1e-4
    assert 0.10 < k["css_mds_large"]["fp64"]["frac"] < 0.12
    assert k["css_mds_large"]["algorithmic_bytes"] == 2600 * 72000.0
    tr = k["css_mds_large"]["traffic"]
    if tr is not None:                                    # profiles/ncu_traffic.json: one 296-window launch, scaled to the chromosome
        assert abs(tr / 2600 - bench.ncu_record("css_mds_large") / 296) < 1.0
        assert 1000 < k["css_mds_large"]["traffic_over_algorithmic"] < 1500
    assert abs(k["css_perm"]["ms_per_chromosome"] - 52.6175) < 1e-3 and k["css_perm"]["scopes_per_chromosome"] == 1
    assert abs(sum(v["share_of_step"] for v in k.values()) - 1.0) < 1e-3
    assert abs(sum(out["kernel_ms_per_chromosome"].values()) - sum(v["ms"] for v in prof.values()) / 16) < 1e-9


def test_bench_headline_kernel_table_follows_its_work_models():
    """bench.py's headline kernel table on the profile of the committed round-2 line (3 steps x 21 chromosome launches): the
    permutation kernel's SURVEY 8(d) gather model against the shared-memory peak, its issue utilisation per scored window (so a
    rank's partial chromosomes at N > 1 stay comparable), and the dense fp64 model of the two MDS kernels."""
    import importlib
    bench = importlib.import_module("bench")
    prof = {"css_pack": {"ms": 1.4613, "launches": 63}, "window_table": {"ms": 1.0317, "launches": 63},
            "css_tridiag": {"ms": 83.4677, "launches": 63}, "css_eigvec": {"ms": 65.8804, "launches": 63},
            "css_observed": {"ms": 8.6592, "launches": 63}, "css_perm": {"ms": 197.657, "launches": 63}}
    css = {"prof": prof, "my_windows": 899997, "my_scored": 815731}
    micro = {"fp64": {"tflops": 33.783}, "smem": {"tb_per_s": 36.992}, "issue": {"gwarp_inst_per_s": 1116.1}}
    k = bench.css_kernel_table(css, micro, 6552.6, 3)
    gath = 815731 * 1000 * (400 + 38) * 8.0
    assert k["css_perm"]["algorithmic_bytes_per_step"] == gath
    assert abs(k["css_perm"]["frac"] - gath / (197.657e-3 / 3) / 36.992e12) < 1e-9 and 1.16 < k["css_perm"]["frac"] < 1.18
    assert abs(k["css_perm"]["ms_per_launch"] - 3.13741) < 1e-4
    iss = k["css_perm"].get("issue")
    if iss is not None:                                   # profiles/ncu_inst.json present
        assert abs(iss["frac"] - iss["warp_inst_per_scored_window"] * 815731 / (197.657e-3 / 3) / 1116.1e9) < 1e-9
        assert 0.55 < iss["frac"] < 0.62
        half = dict(css, my_scored=815731 // 2, prof={n: {"ms": v["ms"] / 2, "launches": v["launches"]} for n, v in prof.items()})
        assert abs(bench.css_kernel_table(half, micro, 6552.6, 3)["css_perm"]["issue"]["frac"] - iss["frac"]) < 1e-3
    fl = 815731 * (6.0 * 1600 + 9.0 * 64000)
    assert abs(k["css_tridiag"]["algorithmic_flops_per_step"] - fl * 2 / 3) < 1.0
    assert abs(k["css_eigvec"]["frac"] - fl / 3 / (65.8804e-3 / 3) / 33.783e12) < 1e-9
    assert abs(sum(v["share_of_step"] for v in k.values()) - 1.0) < 1e-3


def test_bench_finds_the_committed_capture_records_it_quotes():
    """the DRAM traffic and instruction counts bench.py puts beside its live timings come from profiles/ncu_*.json: every
    record the bench asks for is there (a missing one would silently turn `traffic` into null)"""
    import importlib
    bench = importlib.import_module("bench")
    for k in ("css_perm", "css_pack", "css_mds_large", "fet_count", "fet_score"):
        assert bench.ncu_record(k) and bench.ncu_record(k) > 0, k
    assert bench.ncu_record("css_perm", "inst") > 1e9
    assert bench.NCU_LARGE_WINDOWS == 296 and bench.NCU_CSS_SCORED == 38810
