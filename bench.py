#!/usr/bin/env python
"""bench.py — throughput of the two divergence scans on B200, with the reference's CPU path timed beside it.

Headline workload (BASELINE.json configs[2], SURVEY.md 8(d) "C3"): CSS over ONE 450 Mb stickleback-sized synthetic
genome — 21 chromosomes, 1 SNP / 100 bp (4.5 M SNPs), 20+20 diploid individuals, wsize 2500 / wstep 500
(~900 k windows), classical MDS, mcT = mcR = 1000 permutations per window. One step = one pass over the genome.

With N GPUs the genome is split into N CONTIGUOUS window ranges (fpt_b200/sharding.py: chromosomes are dealt in order, the
chromosome a boundary falls into is cut with `partition_windows` / `snp_slice`, each side taking its SNP halo), every rank
scans its range on its own GPU with no exchange, and scores + p are gathered ONCE with NCCL inside the timed region
(`"scaling": "strong"`). Random streams are keyed by the window index, so the gathered result is the single-GPU result.

  value     windows/s with the float64 genotype arrays already resident in HBM (device API, CUDA events, max over ranks)
  e2e       windows/s through the host call (N = 1: the drop-in fpt_css_compute, one call per chromosome as the reference's
            callers make them; N > 1: fpt_css_scan on every rank's range) — host -> device copies of every input from pinned
            host memory and device -> host reads of the results inside the timed region
  kernels   per kernel: live CUDA-event time per launch, share of the step, and algorithmic work / time / the matching
            MEASURED peak (HBM from MEASURED_PEAKS.json; fp64, shared memory, issue slots, u8 tensor core measured by
            profiles/microbench/peaks on this box at the start of the run)
  fet       the FET scan (BASELINE configs[0] geometry) split over the ranks the same way, and BASELINE configs[3]: 100 M
            direct 2x2 tables with coverage <= 500, 100/N M per GPU, scores gathered at the end
  large_cohort  BASELINE configs[4] cohort: 500+500 individuals, 50 kb windows, whole chromosomes of 2600 windows dealt over
            the ranks
  replicas  (N > 1) the round-1 weak-scaling figure: one whole genome per GPU

`--impl reference` times the reference's own pthreads C (oracle/_ref, compiled unmodified from the reference tree)
on the host cores, on a bounded sample of the same workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# ------------------------------------------------------------------------------------------------ workload
CSS = dict(chromosomes=21, length=21_428_500, nsnp=214_285, asize=20, bsize=20, wsize=2500, wstep=500, mds=0,
           mct=1000, mcr=1000, seed0=20261018 + 2)
FET = dict(length=100_000_000, nsnp=1_000_000, asize=20, bsize=20, wsize=2500, wstep=500, perc=0.95, seed0=20261018 + 0)
FET_TABLES = dict(n=100_000_000, chunks=8, lo=20, hi=500, seed0=20261018 + 3)
LARGE = dict(seed0=20261018 + 4, asize=500, bsize=500, chromosomes=8, windows=2600, wsize=50_000, wstep=50_000,
             snps_per_window=167, mcr=1000, cpu_windows=4)
SEED = 20261018
NCU_CSS_SCORED = 38810        # scored windows of the chromosome launch the committed headline captures hold (chromosome 0 of the workload)
NCU_LARGE_WINDOWS = 296      # windows in the one launch the committed large-cohort ncu captures hold (profiles/capture.sh: probe_large_cohort.py 296)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def micro_peaks():
    """fp64 / shared-memory / issue / u8-tensor peaks measured on THIS box by profiles/microbench/peaks (a few seconds);
    the committed copy of an earlier run (profiles/r2_peaks.json) is the fallback and is named as such"""
    exe = os.path.join(ROOT, "profiles", "microbench", "peaks")
    try:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        if r.returncode == 0:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            d["source"] = "profiles/microbench/peaks run on this box at the start of this bench"
            return d
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "r2_peaks.json")) as f:
            d = json.load(f)
        d["source"] = "profiles/r2_peaks.json (committed copy of an earlier run; the live microbenchmark did not run)"
        return d
    except Exception:
        return None


def ncu_record(kernel, what="traffic"):
    """per-launch DRAM bytes (or warp instructions, what="inst") of `kernel` from the committed ncu capture"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_%s.json" % what)) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """samples SM clock and throttle reasons every 100 ms while the timed region runs"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.th is not None:
            self.th.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ helpers
def pinned_copy(arr):
    import torch
    t = torch.empty(arr.shape, dtype=torch.from_numpy(arr[:0]).dtype).pin_memory()
    v = t.numpy()
    v[...] = arr
    return t, v


class Ctx:
    """one rank's view of the job"""

    def __init__(self, args, rank, world, dist):
        import torch
        self.args, self.rank, self.world, self.dist, self.torch = args, rank, world, dist, torch
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.current_stream()
        self.sp = C.c_void_p(self.stream.cuda_stream)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def timed(self, step, nsteps, flush=None):
        """K steps between barriers, CUDA events on the launching stream, max over ranks -> ms for all K steps"""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if flush is None:
            self.barrier()
            e0.record(self.stream)
            for _ in range(nsteps):
                step()
            e1.record(self.stream)
            self.barrier()
            ms = e0.elapsed_time(e1)
        else:                                            # L2 flushed between timed iterations
            ms = 0.0
            for _ in range(nsteps):
                flush.zero_()
                self.barrier()
                e0.record(self.stream)
                step()
                e1.record(self.stream)
                self.barrier()
                ms += e0.elapsed_time(e1)
        return self.max_over_ranks(ms)[0]


def profile_json(lib):
    buf = C.create_string_buffer(16384)
    lib.fpt_profile_summary(buf, 16384)
    return json.loads(buf.value.decode() or "{}")


def genome_pieces(nchrom, nout, rank, world):
    """the contiguous window range of the whole genome owned by `rank`, as [(chromosome, window_begin, window_end)]"""
    from fpt_b200.sharding import partition_windows
    gb, ge = partition_windows(nchrom * nout, world)[rank]
    out = []
    for c in range(nchrom):
        b, e = max(gb, c * nout), min(ge, (c + 1) * nout)
        if e > b:
            out.append((c, b - c * nout, e - c * nout))
    return out, (gb, ge)


# ------------------------------------------------------------------------------------------------ B200 arm: CSS
def bench_css(lib_mod, cx, full_genome_pass):
    import torch
    import fpt_b200.synth as synth
    from fpt_b200._lib import Genotypes, ScanRange, check
    from fpt_b200.sharding import partition_windows, snp_slice
    args, rank, world, dev, sp = cx.args, cx.rank, cx.world, cx.dev, cx.sp
    lib = lib_mod.load()
    asize, bsize, m = CSS["asize"], CSS["bsize"], CSS["asize"] + CSS["bsize"]
    regend, wsize, wstep, nsnp = CSS["length"], CSS["wsize"], CSS["wstep"], CSS["nsnp"]
    nout, nchrom = regend // wstep, CSS["chromosomes"]
    total = nchrom * nout
    ranges = partition_windows(total, world)
    pieces, (gb, ge) = genome_pieces(nchrom, nout, rank, world)
    need = sorted(set(c for c, _, _ in pieces)) if not full_genome_pass else list(range(nchrom))
    chroms = {c: synth.chromosome_fast(CSS["seed0"] + c, regend, nsnp, asize, bsize, wstep=wstep) for c in need}
    # reference layout per chromosome: float64 values + repeated int32 positions (host, pinned) and the values resident in HBM
    host, resident = {}, {}
    for c in need:
        av, bv, apos, bpos = synth.reference_layout(chroms[c])
        host[c] = [pinned_copy(x) for x in (av, bv, apos, bpos)] + [pinned_copy(chroms[c]["pos"])]
        resident[c] = (host[c][0][0].to(dev), host[c][1][0].to(dev), torch.from_numpy(chroms[c]["pos"]).to(dev))
    planes = torch.empty(lib.fpt_dev_css_planes_bytes(nsnp, m) // 4 + 64, dtype=torch.int32, device=dev)
    wl = torch.empty(nout, dtype=torch.int32, device=dev)
    wr = torch.empty(nout, dtype=torch.int32, device=dev)
    mx = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.fpt_dev_css_workspace_bytes(m, nout, CSS["mds"])
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    longest = max(e - b for b, e in ranges)
    local = torch.zeros(2 * longest, dtype=torch.float64, device=dev)        # my scores, then my p (padded to the longest range)
    status = torch.zeros(longest, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * 2 * longest, dtype=torch.float64, device=dev) if world > 1 else None

    def piece_plan(plist):
        """per piece: SNP slice with halo, 32-SNP aligned start (whole bit-plane words), window range, output offset"""
        plan, off = [], 0
        for (c, wb, we) in plist:
            pos = chroms[c]["pos"]
            lo, hi = snp_slice(pos, wb, we, wsize, wstep)
            r = ScanRange()
            r.regend, r.wsize, r.wstep, r.semantics, r.window_begin, r.window_end, r.seed = regend, wsize, wstep, 0, wb, we, SEED
            plan.append((c, lo, hi, r, off, we - wb))
            off += we - wb
        return plan

    plan = piece_plan(pieces)

    def run_pieces(plan_, out_s, out_p, out_st):
        for (c, lo, hi, r, off, nw) in plan_:
            da, db, dpos = resident[c]
            ns = hi - lo
            check(lib.fpt_dev_css_pack_f64(da.data_ptr() + lo * asize * 8, db.data_ptr() + lo * bsize * 8, ns, asize, bsize, planes.data_ptr(), sp))
            mx.zero_()
            check(lib.fpt_dev_window_table(dpos.data_ptr() + lo * 4, ns, C.byref(r), wl.data_ptr(), wr.data_ptr(), mx.data_ptr(), sp))
            check(lib.fpt_dev_css_windows(planes.data_ptr(), None, asize, bsize, wl.data_ptr(), wr.data_ptr(), C.byref(r),
                                          CSS["mct"], CSS["mcr"], CSS["mds"], ws.data_ptr(), ws_bytes, out_s.data_ptr() + off * 8,
                                          out_p.data_ptr() + off * 8, out_st.data_ptr() + off, None, sp))

    def step_resident():
        status.zero_()
        run_pieces(plan, local, local[longest:], status)
        if world > 1:                                    # the only exchange: results gathered once at the end
            cx.dist.all_gather_into_tensor(gathered, local)

    launches_per_step = len(plan) * 6 + (1 if world > 1 else 0)   # pack, window table, tridiagonalisation, eigenvectors, observed scores, permutations (+ gather)

    def e2e_piece(c, lo, hi, r, nw):
        """host call on this rank's range of chromosome c: float64 host arrays (pinned) in, host arrays out"""
        av, bv, apos, bpos, hpos = host[c]
        s, p = np.zeros(nw), np.zeros(nw)
        if world == 1:                                   # the reference's own call: the drop-in, whole chromosome
            check(lib.fpt_css_compute(av[1].ctypes.data, bv[1].ctypes.data, apos[1].ctypes.data, bpos[1].ctypes.data, 0, regend,
                                      wsize, wstep, av[1].size, bv[1].size, CSS["mct"], CSS["mcr"], 0, CSS["mds"],
                                      s.ctypes.data, p.ctypes.data))
        else:
            g = Genotypes()
            g.avals, g.bvals = av[1].ctypes.data + lo * asize * 8, bv[1].ctypes.data + lo * bsize * 8
            g.pos, g.nsnp, g.asize, g.bsize = hpos[1].ctypes.data + lo * 4, hi - lo, asize, bsize
            check(lib.fpt_css_scan(C.byref(g), C.byref(r), CSS["mct"], CSS["mcr"], 0, CSS["mds"], s.ctypes.data, p.ctypes.data, None, None))
        return s, p

    host_local = torch.zeros(2 * longest, dtype=torch.float64).pin_memory()

    def step_e2e():
        scored = 0
        hv = host_local.numpy()
        for (c, lo, hi, r, off, nw) in plan:
            s, p = e2e_piece(c, lo, hi, r, nw)
            hv[off:off + nw] = s
            hv[longest + off:longest + off + nw] = p
            scored += int(np.count_nonzero(p))
        if world > 1:                                    # gather over NVLink, then every rank holds the genome's results on the host
            local.copy_(host_local, non_blocking=True)
            cx.dist.all_gather_into_tensor(gathered, local)
            _ = gathered.cpu()
        return scored

    for _ in range(args.warmup):
        step_resident()
    cx.barrier()
    lib.fpt_profile_enable(1)
    profile_json(lib)                                    # clear
    with ClockSampler(torch.cuda.current_device()) as clk:
        ms = cx.timed(step_resident, args.steps)
    prof = profile_json(lib)
    lib.fpt_profile_enable(0)
    my_scored = int((status == 2).sum().item())
    scored = int(cx.sum_over_ranks(my_scored)[0])
    rechecks = int(lib.fpt_css_perm_rechecks())          # exact re-scorings since the library was loaded (all steps so far)
    sample = None
    if rank == 0:
        n0 = plan[0][5]
        sample = (chroms[0], host[0], local[:n0].cpu().numpy(), local[longest:longest + n0].cpu().numpy())

    # the other two MDS variants of BASELINE configs[2], on chromosome 0 only (rank 0, device-resident, warm-up + one timed pass each)
    variants = {}
    if rank == 0:
        da, db, dpos = resident[0]
        r = ScanRange()
        r.regend, r.wsize, r.wstep, r.semantics, r.window_begin, r.window_end, r.seed = regend, wsize, wstep, 0, 0, nout, SEED
        check(lib.fpt_dev_css_pack_f64(da.data_ptr(), db.data_ptr(), nsnp, asize, bsize, planes.data_ptr(), sp))
        mx.zero_()
        check(lib.fpt_dev_window_table(dpos.data_ptr(), nsnp, C.byref(r), wl.data_ptr(), wr.data_ptr(), mx.data_ptr(), sp))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        from fpt_b200._lib import CssProbes
        for mds_v in (1, 2):
            nruns = 4 if mds_v == 1 else 1
            wsb = lib.fpt_dev_css_workspace_bytes(m, nout, mds_v)
            wsv = torch.empty(wsb, dtype=torch.uint8, device=dev)
            sv, pv, stv = (torch.zeros(nout, dtype=torch.float64, device=dev), torch.zeros(nout, dtype=torch.float64, device=dev),
                           torch.zeros(nout, dtype=torch.uint8, device=dev))
            iters = torch.zeros(nout * nruns, dtype=torch.int32, device=dev)
            pr = CssProbes()
            pr.smacof_iters = iters.data_ptr()
            lib.fpt_profile_enable(1)
            for it in range(2):
                stv.zero_()
                torch.cuda.synchronize()
                profile_json(lib)
                e0.record(cx.stream)
                check(lib.fpt_dev_css_windows(planes.data_ptr(), None, asize, bsize, wl.data_ptr(), wr.data_ptr(), C.byref(r),
                                              CSS["mct"], CSS["mcr"], mds_v, wsv.data_ptr(), wsb, sv.data_ptr(), pv.data_ptr(),
                                              stv.data_ptr(), C.byref(pr), sp))
                e1.record(cx.stream)
                torch.cuda.synchronize()
            vprof = profile_json(lib)
            lib.fpt_profile_enable(0)
            nsc = int((stv == 2).sum().item())
            tot_iters = int(iters.view(nout, nruns)[stv == 2].sum().item())
            variants["mds%d" % mds_v] = {"windows_per_s": nout / (e0.elapsed_time(e1) * 1e-3), "ms_per_chromosome": e0.elapsed_time(e1),
                                         "windows_scored": nsc, "smacof_iterations_total": tot_iters,
                                         "kernel_ms": {k: v["ms"] for k, v in vprof.items()}}
            del wsv
            # the same chromosome end to end through the drop-in (pinned host float64 arrays in, host results out)
            hav, hbv, hapos, hbpos, _ = host[0]
            ts, hs = [], None
            for it in range(2):
                hs, hp = np.zeros(nout), np.zeros(nout)
                t0 = time.perf_counter()
                check(lib.fpt_css_compute(hav[1].ctypes.data, hbv[1].ctypes.data, hapos[1].ctypes.data, hbpos[1].ctypes.data, 0, regend,
                                          wsize, wstep, hav[1].size, hbv[1].size, CSS["mct"], CSS["mcr"], 0, mds_v, hs.ctypes.data, hp.ctypes.data))
                ts.append(time.perf_counter() - t0)
            variants["mds%d" % mds_v]["e2e"] = {"value": nout / ts[-1], "unit": "windows/s", "ms_per_chromosome": ts[-1] * 1e3,
                                                "api": "fpt_css_compute (drop-in, pinned host float64 arrays), chromosome 0",
                                                "h2d_bytes": int(hav[1].nbytes + hbv[1].nbytes + hapos[1].nbytes + hbpos[1].nbytes), "d2h_bytes": 16 * nout}
            variants["mds%d" % mds_v]["_scores"] = hs

    # end to end through the host call
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        scored_e2e = step_e2e()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    ms_e2e = cx.max_over_ranks(t_e2e * 1000.0)[0]
    scored_e2e = int(cx.sum_over_ranks(scored_e2e)[0])
    h2d = sum((hi - lo) * (m * 8 + 4) for (c, lo, hi, r, off, nw) in plan)
    d2h = sum(nw * 17 for (c, lo, hi, r, off, nw) in plan)
    h2d, d2h = [int(x) for x in cx.sum_over_ranks(h2d, d2h)]

    # pageable host memory (what HyperBrowser hands over): the drop-in on chromosome 0 with plain numpy arrays, rank 0, N = 1 only
    pageable = None
    if world == 1:
        av, bv, apos, bpos = synth.reference_layout(chroms[0])
        ts = []
        for it in range(3):
            s, p = np.zeros(nout), np.zeros(nout)
            t0 = time.perf_counter()
            check(lib.fpt_css_compute(av.ctypes.data, bv.ctypes.data, apos.ctypes.data, bpos.ctypes.data, 0, regend, wsize, wstep, av.size,
                                      bv.size, CSS["mct"], CSS["mcr"], 0, CSS["mds"], s.ctypes.data, p.ctypes.data))
            ts.append(time.perf_counter() - t0)
        pageable = {"value": nout / min(ts[1:]), "unit": "windows/s", "api": "fpt_css_compute, pageable numpy arrays, chromosome 0",
                    "ms_per_chromosome": min(ts[1:]) * 1e3}

    # secondary: one whole genome per GPU (round-1 weak-scaling replicas)
    replicas = None
    if world > 1 and full_genome_pass:
        plan_full = piece_plan([(c, 0, nout) for c in range(nchrom)])
        fs = torch.zeros(total, dtype=torch.float64, device=dev)
        fp_ = torch.zeros(total, dtype=torch.float64, device=dev)
        fst = torch.zeros(total, dtype=torch.uint8, device=dev)
        run_pieces(plan_full, fs, fp_, fst)
        ms_rep = cx.timed(lambda: run_pieces(plan_full, fs, fp_, fst), 2)
        replicas = {"value": world * total / (ms_rep / 2 * 1e-3), "unit": "windows/s", "scaling": "weak", "ms_per_step": ms_rep / 2,
                    "config": "one whole 450 Mb genome per GPU, no exchange (the round-1 figure)"}
    return dict(ms=ms, ms_e2e=ms_e2e, total_windows=total, scored=scored, scored_e2e=scored_e2e, prof=prof, my_windows=ge - gb,
                launches=launches_per_step * args.steps, h2d=h2d, d2h=d2h, rechecks=rechecks, variants=variants, clocks=clk.summary(),
                sample=sample, pageable=pageable, replicas=replicas, pieces=len(plan), my_scored=my_scored)


# ------------------------------------------------------------------------------------------------ B200 arm: FET
def bench_fet(lib_mod, cx):
    import torch
    import fpt_b200.synth as synth
    from fpt_b200._lib import Genotypes, ScanRange, check
    from fpt_b200.sharding import partition_windows, snp_slice
    args, rank, world, dev, sp = cx.args, cx.rank, cx.world, cx.dev, cx.sp
    lib = lib_mod.load()
    ch = synth.chromosome_fast(FET["seed0"], FET["length"], FET["nsnp"], FET["asize"], FET["bsize"], wstep=FET["wstep"])
    nsnp, asize, bsize = FET["nsnp"], FET["asize"], FET["bsize"]
    regend, wsize, wstep = FET["length"], FET["wsize"], FET["wstep"]
    nout = regend // wstep
    ranges = partition_windows(nout, world)
    wb, we = ranges[rank]
    lo, hi = snp_slice(ch["pos"], wb, we, wsize, wstep)
    ns, nw = hi - lo, we - wb
    sub = {"pos": ch["pos"][lo:hi], "acodes": ch["acodes"][lo * asize:hi * asize], "bcodes": ch["bcodes"][lo * bsize:hi * bsize],
           "asize": asize, "bsize": bsize}
    av, bv, apos, bpos = synth.reference_layout(sub)
    hp = [pinned_copy(x) for x in (av, bv, apos, bpos, sub["pos"], sub["acodes"], sub["bcodes"])]
    da, db = hp[0][0].to(dev), hp[1][0].to(dev)
    dpos = hp[4][0].to(dev)
    tab = torch.empty(ns * 4, dtype=torch.int32, device=dev)
    snp = torch.empty(ns, dtype=torch.float64, device=dev)
    wl = torch.empty(nw, dtype=torch.int32, device=dev)
    wr = torch.empty(nw, dtype=torch.int32, device=dev)
    mx = torch.zeros(1, dtype=torch.int32, device=dev)
    longest = max(e - b for b, e in ranges)
    local = torch.zeros(2 * longest, dtype=torch.float64, device=dev)
    ofl = torch.zeros(nw, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * 2 * longest, dtype=torch.float64, device=dev) if world > 1 else None
    r = ScanRange()
    r.regend, r.wsize, r.wstep, r.semantics, r.window_begin, r.window_end, r.seed = regend, wsize, wstep, 0, wb, we, SEED
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > L2 (126 MB)

    def step():
        check(lib.fpt_dev_fet_count_f64(da.data_ptr(), db.data_ptr(), ns, asize, bsize, tab.data_ptr(), sp))
        check(lib.fpt_dev_fet_score(tab.data_ptr(), ns, asize + bsize, 0, snp.data_ptr(), sp))
        mx.zero_()
        check(lib.fpt_dev_window_table(dpos.data_ptr(), ns, C.byref(r), wl.data_ptr(), wr.data_ptr(), mx.data_ptr(), sp))
        max_npos = int(mx.item())                        # the one host read the scan needs (sizes shared memory)
        ofl.zero_()
        check(lib.fpt_dev_fet_windows(snp.data_ptr(), wl.data_ptr(), wr.data_ptr(), C.byref(r), max_npos, FET["perc"], None,
                                      local.data_ptr(), local.data_ptr() + longest * 8, ofl.data_ptr(), sp))
        if world > 1:
            cx.dist.all_gather_into_tensor(gathered, local)

    for _ in range(args.warmup):
        step()
    cx.barrier()
    lib.fpt_profile_enable(1)
    profile_json(lib)
    ms = cx.timed(step, args.steps, flush=flush) / args.steps
    prof = profile_json(lib)
    lib.fpt_profile_enable(0)

    def host_call(which):
        s, d = np.zeros(nw), np.zeros(nw)
        if which == "dropin":                            # N = 1: the reference's call
            check(lib.fpt_fet_compute(hp[0][1].ctypes.data, hp[1][1].ctypes.data, hp[2][1].ctypes.data, hp[3][1].ctypes.data, 0,
                                      regend, wsize, wstep, hp[0][1].size, hp[1][1].size, FET["perc"], s.ctypes.data, d.ctypes.data))
        else:
            g = Genotypes()
            if which == "f64":
                g.avals, g.bvals = hp[0][1].ctypes.data, hp[1][1].ctypes.data
            else:
                g.acodes, g.bcodes = hp[5][1].ctypes.data, hp[6][1].ctypes.data
            g.pos, g.nsnp, g.asize, g.bsize = hp[4][1].ctypes.data, ns, asize, bsize
            check(lib.fpt_fet_scan(C.byref(g), C.byref(r), FET["perc"], s.ctypes.data, d.ctypes.data, None))
        return s, d

    host_local = torch.zeros(2 * longest, dtype=torch.float64).pin_memory()

    def e2e(which):
        s, d = host_call(which)
        if world > 1:
            hv = host_local.numpy()
            hv[:nw] = s
            hv[longest:longest + nw] = d
            local.copy_(host_local, non_blocking=True)
            cx.dist.all_gather_into_tensor(gathered, local)
            _ = gathered.cpu()
        return s

    def time_e2e(which):
        for _ in range(2):
            e2e(which)
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            s_host = e2e(which)
        torch.cuda.synchronize()
        return cx.max_over_ranks((time.perf_counter() - t0) * 1000.0 / args.steps)[0], s_host

    ms_e2e, s_host = time_e2e("dropin" if world == 1 else "f64")
    ms_e2e_i8, _ = time_e2e("i8")
    pageable = None
    if world == 1:
        pav, pbv, papos, pbpos = (np.array(hp[k][1]) for k in range(4))         # plain (pageable) numpy copies
        ts = []
        for it in range(3):
            s, d = np.zeros(nw), np.zeros(nw)
            t0 = time.perf_counter()
            check(lib.fpt_fet_compute(pav.ctypes.data, pbv.ctypes.data, papos.ctypes.data, pbpos.ctypes.data, 0, regend, wsize, wstep,
                                      pav.size, pbv.size, FET["perc"], s.ctypes.data, d.ctypes.data))
            ts.append(time.perf_counter() - t0)
        pageable = {"value": nsnp / min(ts[1:]), "unit": "SNPs/s", "api": "fpt_fet_compute, pageable numpy arrays"}
        del pav, pbv, papos, pbpos
    out = {
        "metric": "fet_snps_per_sec", "unit": "SNPs/s", "value": nsnp / (ms * 1e-3), "ms_per_step": ms, "n_gpus": world,
        "scaling": "strong",
        "config": {"workload": "FET scan, 1 chromosome 100 Mb, 1M SNPs, 20+20, wsize 2500 / wstep 500, perc 0.95 (BASELINE configs[0] "
                               "geometry), windows split into %d contiguous ranges with SNP halos, scores + sigma gathered once; "
                               "L2 flushed between steps" % world, "windows": nout},
        "dtype": "f64",
        "e2e": {"value": nsnp / (ms_e2e * 1e-3), "unit": "SNPs/s", "h2d_bytes_per_step": nsnp * ((asize + bsize) * 8 + 4),
                "d2h_bytes_per_step": nout * 17,
                "api": "fpt_fet_compute (drop-in, host float64 arrays, pinned)" if world == 1 else "fpt_fet_scan per rank (host float64, pinned) + NCCL gather"},
        "e2e_int8": {"value": nsnp / (ms_e2e_i8 * 1e-3), "unit": "SNPs/s", "h2d_bytes_per_step": nsnp * ((asize + bsize) + 4),
                     "d2h_bytes_per_step": nout * 17, "api": "fpt_fet_scan (host int8 codes, pinned): 8x fewer PCIe bytes than the drop-in layout"},
        "e2e_pageable": pageable,
        "gpu_launches": (4 + (1 if world > 1 else 0)) * args.steps,
        "_prof": prof, "_ns": ns, "_nw": nw,
    }
    return out, (ch, s_host, wb, we)


def tables_chunk(torch, dev, k, n):
    """chunk k of the configs[3] tables: row sums U{lo..hi}, a ~ Bin(n1, f), c ~ Bin(n2, f_B); the same data
    whatever the number of ranks"""
    g = torch.Generator(device=dev)
    g.manual_seed(FET_TABLES["seed0"] + 1000 * k)
    n1 = torch.randint(FET_TABLES["lo"], FET_TABLES["hi"] + 1, (n,), device=dev, generator=g).double()
    n2 = torch.randint(FET_TABLES["lo"], FET_TABLES["hi"] + 1, (n,), device=dev, generator=g).double()
    f = torch.rand(n, device=dev, generator=g, dtype=torch.float64).clamp(0.02, 0.98)
    fb = (f + 0.05 * torch.randn(n, device=dev, generator=g, dtype=torch.float64)).clamp(0.01, 0.99)
    a = torch.binomial(n1, f, generator=g)
    c = torch.binomial(n2, fb, generator=g)
    return torch.stack([a, n1 - a, c, n2 - c], dim=1).to(torch.int32).contiguous()


def walk_lengths(T):
    """per table: a0 = first-tail length (the minimum cell) and T2 = terms of the opposite tail the reference's walk visits
    (tables on the far side whose probability is below the observed one; cFisher.c:405-455) — SURVEY 8(d) K2 work model"""
    from scipy.special import gammaln
    T = T.astype(np.int64)
    a, b, c, d = T[:, 0], T[:, 1], T[:, 2], T[:, 3]
    R1, C1, N = a + b, a + c, a + b + c + d
    lo_, hi_ = np.maximum(0, R1 + C1 - N), np.minimum(R1, C1)
    a0 = np.minimum(np.minimum(a, b), np.minimum(c, d))
    width = int((hi_ - lo_).max()) + 1
    x = lo_[:, None] + np.arange(width)[None, :]
    valid = x <= hi_[:, None]
    xa = np.where(valid, x, lo_[:, None])

    def lpmf(xx):
        return -(gammaln(xx + 1) + gammaln(R1[:, None] - xx + 1) + gammaln(C1[:, None] - xx + 1) + gammaln(N[:, None] - R1[:, None] - C1[:, None] + xx + 1))
    lp = lpmf(xa)
    lobs = lpmf(a[:, None])
    # the observed side: the tail that runs towards the extreme where the minimum cell reaches 0
    towards_low = (a == a0) | (d == a0)                  # a (or d) shrinks -> x decreases
    far = np.where(towards_low[:, None], xa > a[:, None], xa < a[:, None])
    t2 = (valid & far & (lp < lobs - 1e-12)).sum(axis=1)
    sym = (R1 == (N - R1)) | (C1 == (N - C1))            # cFisher.c:430: P doubled, no second walk
    t2 = np.where(sym, 0, t2)
    return a0, t2


def bench_fet_tables(lib_mod, cx):
    import torch
    from fpt_b200._lib import check
    args, rank, world, dev, sp = cx.args, cx.rank, cx.world, cx.dev, cx.sp
    lib = lib_mod.load()
    nchunks = FET_TABLES["chunks"]
    per = FET_TABLES["n"] // nchunks
    mine = [k for k in range(nchunks) if k * world // nchunks == rank] if world <= nchunks else ([rank] if rank < nchunks else [])
    T = torch.cat([tables_chunk(torch, dev, k, per) for k in mine]) if mine else torch.zeros((0, 4), dtype=torch.int32, device=dev)
    n_local = T.shape[0]
    n_total = per * nchunks
    longest = per * max(1, nchunks // min(world, nchunks))
    outp = torch.zeros(longest, dtype=torch.float64, device=dev)
    gathered = torch.empty(world * longest, dtype=torch.float64, device=dev) if world > 1 else None

    def step():
        if n_local:
            check(lib.fpt_dev_fet_score(T.data_ptr(), n_local, 1000, 0, outp.data_ptr(), sp))
        if world > 1:
            cx.dist.all_gather_into_tensor(gathered, outp)

    for _ in range(args.warmup):
        step()
    cx.barrier()
    lib.fpt_profile_enable(1)
    profile_json(lib)
    ms = cx.timed(step, args.steps) / args.steps         # inputs (>= 200 MB per GPU) exceed L2, no flush needed
    prof = profile_json(lib)
    lib.fpt_profile_enable(0)
    k_ms = prof.get("fet_score", {}).get("ms", 0.0) / max(1, prof.get("fet_score", {}).get("launches", 1))
    k_ms = cx.max_over_ranks(k_ms)[0]
    # end to end: host int32 tables (pinned) -> host float64 scores
    e2e_ms = None
    if n_local:
        hT = torch.empty((n_local, 4), dtype=torch.int32).pin_memory()
        hT.copy_(T)
        ho = np.zeros(n_local)
        for _ in range(1):
            check(lib.fpt_fet_tables(hT.data_ptr(), n_local, 0, ho.ctypes.data))
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps - 1)):
        if n_local:
            check(lib.fpt_fet_tables(hT.data_ptr(), n_local, 0, ho.ctypes.data))
        if world > 1:
            outp[:n_local].copy_(torch.from_numpy(ho), non_blocking=True)
            cx.dist.all_gather_into_tensor(gathered, outp)
            _ = gathered.cpu()
    torch.cuda.synchronize()
    e2e_ms = cx.max_over_ranks((time.perf_counter() - t0) * 1000.0 / max(1, args.steps - 1))[0]
    # work model on a sample of this rank's tables (rank 0 reports)
    work = None
    if rank == 0 and n_local:
        smp = T[:: max(1, n_local // 131072)][:131072].cpu().numpy()
        a0, t2 = walk_lengths(smp)
        Tw = (a0 + t2).astype(np.float64)
        work = {"sample_tables": int(smp.shape[0]), "mean_first_tail": float(a0.mean()), "mean_second_tail": float(t2.mean()),
                "mean_T": float(Tw.mean()), "p99_T": float(np.percentile(Tw, 99)), "max_T": float(Tw.max()),
                "sum_T_scaled_to_all_tables": float(Tw.mean() * n_total),
                "warp_divergence_factor": float(np.mean([Tw[i:i + 32].max() for i in range(0, len(Tw) - 31, 32)]) / max(Tw.mean(), 1e-9)),
                "note": "T = a0 + T2 per table: steps of the reference's two-tailed walk (cFisher.c:405-455); warp_divergence_factor = "
                        "mean over warps of the longest walk in the warp / mean walk (1 = no divergence loss with one thread per table)"}
    return {"metric": "fet_snps_per_sec", "unit": "SNPs/s", "value": n_total / (ms * 1e-3), "ms_per_step": ms, "n_gpus": world,
            "scaling": "strong", "dtype": "f64",
            "config": {"workload": "BASELINE configs[3]: %d M direct 2x2 tables, row sums U{20..500}, log-space arithmetic, %s tables per GPU, "
                                   "scores gathered with NCCL at the end (inside the timed region); inputs exceed L2" % (
                                       n_total // 1_000_000, "%.1f M" % (n_total / world / 1e6)), "tables": n_total},
            "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": "SNPs/s", "h2d_bytes_per_step": n_total * 16, "d2h_bytes_per_step": n_total * 8,
                    "api": "fpt_fet_tables (host int32 tables, pinned)" + (" per rank + NCCL gather" if world > 1 else "")},
            "gpu_launches": args.steps, "_kernel_ms": k_ms, "_work": work, "_n_local_max": longest}


# ------------------------------------------------------------------------------------------------ B200 arm: large cohort
def large_chromosome(torch, dev, nwin):
    """one configs[4] chromosome generated on the device (a host generator needs minutes for 434 M genotypes): positions uniform
    and sorted, f ~ U(0.02, 0.98) per SNP shared by both populations, genotype = number of minor alleles of two draws,
    codes 3 / 0 / -3, 2 % missing (-128)"""
    g = torch.Generator(device=dev)
    g.manual_seed(LARGE["seed0"])
    nsnp = LARGE["snps_per_window"] * nwin
    regend = LARGE["wstep"] * nwin
    pos = torch.sort(torch.randperm(regend, device=dev, generator=g)[:nsnp]).values.to(torch.int32)
    f = torch.rand(nsnp, device=dev, generator=g) * 0.96 + 0.02
    out = []
    for size in (LARGE["asize"], LARGE["bsize"]):
        u = torch.rand((nsnp, size, 2), device=dev, generator=g)
        minor = (u < f[:, None, None]).sum(dim=2)
        codes = (3 - 3 * minor).to(torch.int8)
        miss = torch.rand((nsnp, size), device=dev, generator=g) < 0.02
        codes[miss] = -128
        out.append(codes.reshape(-1).contiguous())
    return pos.contiguous(), out[0], out[1], regend, nsnp


def bench_large_cohort(lib_mod, cx):
    """BASELINE configs[4] cohort (500+500 individuals, 50 kb windows, ~167 SNPs per window): whole chromosomes of 2600 windows,
    dealt over the ranks (chromosome k goes to rank k % N); one synthetic chromosome is reused with a different random-stream
    seed per chromosome. Kernels: bit-plane packing, window table, Lanczos classical MDS, observed scores (a warp per window),
    permutation test with the between-group sums of 128 permutations as one tcgen05 u8 contraction in tensor memory."""
    import torch
    import fpt_b200.api as api
    from fpt_b200._lib import ScanRange, check
    args, rank, world, dev, sp = cx.args, cx.rank, cx.world, cx.dev, cx.sp
    lib = lib_mod.load()
    nwin = LARGE["windows"] if not args.small else 8
    nchrom = LARGE["chromosomes"] if not args.small else 2
    asize, bsize = LARGE["asize"], LARGE["bsize"]
    m = asize + bsize
    dpos, da, db, regend, nsnp = large_chromosome(torch, dev, nwin)
    mine = [k for k in range(nchrom) if k % world == rank]
    planes = torch.empty(lib.fpt_dev_css_planes_bytes(nsnp, m) // 4 + 64, dtype=torch.int32, device=dev)
    wl = torch.empty(nwin, dtype=torch.int32, device=dev)
    wr = torch.empty(nwin, dtype=torch.int32, device=dev)
    mx = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.fpt_dev_css_workspace_bytes(m, nwin, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    per_rank = -(-nchrom // world)
    local = torch.zeros(2 * per_rank * nwin, dtype=torch.float64, device=dev)
    status = torch.zeros(per_rank * nwin, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * local.numel(), dtype=torch.float64, device=dev) if world > 1 else None

    def step():
        status.zero_()
        for i, k in enumerate(mine):
            r = ScanRange()
            r.regend, r.wsize, r.wstep, r.semantics, r.window_begin, r.window_end, r.seed = regend, LARGE["wsize"], LARGE["wstep"], 0, 0, nwin, SEED + k
            check(lib.fpt_dev_css_pack_i8(da.data_ptr(), db.data_ptr(), nsnp, asize, bsize, planes.data_ptr(), sp))
            mx.zero_()
            check(lib.fpt_dev_window_table(dpos.data_ptr(), nsnp, C.byref(r), wl.data_ptr(), wr.data_ptr(), mx.data_ptr(), sp))
            check(lib.fpt_dev_css_windows(planes.data_ptr(), None, asize, bsize, wl.data_ptr(), wr.data_ptr(), C.byref(r), LARGE["mcr"],
                                          LARGE["mcr"], 0, ws.data_ptr(), ws_bytes, local.data_ptr() + i * nwin * 8,
                                          local.data_ptr() + (per_rank + i) * nwin * 8, status.data_ptr() + i * nwin, None, sp))
        if world > 1:
            cx.dist.all_gather_into_tensor(gathered, local)

    step()                                               # warm-up (first call pays the scratch allocation)
    cx.barrier()
    lib.fpt_profile_enable(1)
    profile_json(lib)
    nsteps = 1 if args.small else max(1, min(args.steps, 2))
    ph = (C.c_ulonglong * 8)()
    lib.fpt_debug_lanczos_phases(ph)                     # reset the phase / step counters
    ms = cx.timed(step, nsteps) / nsteps
    prof = profile_json(lib)
    lib.fpt_profile_enable(0)
    lanczos_steps = None
    if lib.fpt_debug_lanczos_phases(ph) == 0 and len(mine):
        lanczos_steps = float(ph[7]) / (nsteps * len(mine) * nwin)      # mean Lanczos steps per window on this rank
    scored = int(cx.sum_over_ranks(int((status == 2).sum().item()))[0])
    # end to end: host int8 codes (pinned) through fpt_css_scan, one call per chromosome
    hpos, ha, hb = (pinned_copy(x.cpu().numpy()) for x in (dpos, da, db))
    t_e2e = None
    for it in range(2):
        cx.barrier()
        t0 = time.perf_counter()
        for k in mine:
            s, p, w = api.css_scan(ha[1], hb[1], hpos[1], asize, bsize, regend, LARGE["wsize"], LARGE["wstep"], LARGE["mcr"], LARGE["mcr"], mds=0, seed=SEED + k)
        torch.cuda.synchronize()
        t_e2e = cx.max_over_ranks(time.perf_counter() - t0)[0]
    first = None
    if rank == 0:
        first = (hpos[1], ha[1], hb[1], local[:nwin].cpu().numpy(), local[per_rank * nwin:per_rank * nwin + nwin].cpu().numpy(), regend)
    return {"metric": "css_windows_per_sec_1000perms", "unit": "windows/s", "value": nchrom * nwin / (ms * 1e-3), "ms_per_step": ms,
            "n_gpus": world, "scaling": "strong", "dtype": "f64",
            "e2e": {"value": nchrom * nwin / t_e2e, "unit": "windows/s", "api": "fpt_css_scan (host int8 codes, pinned), one call per chromosome",
                    "h2d_bytes_per_step": int(nchrom * (nsnp * m + nsnp * 4)), "d2h_bytes_per_step": int(nchrom * nwin * 17)},
            "config": {"workload": "CSS scan, BASELINE configs[4] cohort: %d chromosomes x %d windows (%d of the 60 000 windows of the 3 Gb genome), 500+500 "
                                   "individuals, 50 kb windows, ~%d SNPs per window, classical MDS, mcT=mcR=1000; chromosome k on rank k %% N, "
                                   "results gathered once" % (nchrom, nwin, nchrom * nwin, LARGE["snps_per_window"]),
                       "windows_scored": scored},
            # kernels launched on this rank per step: one per profile scope, and the Lanczos scope holds two (the arithmetic-product
            # kernel, then the general one for the windows it left pending)
            "gpu_launches": (sum(v["launches"] for v in prof.values()) + prof.get("css_mds_large", {}).get("launches", 0)) // max(1, nsteps) * nsteps,
            "_prof": prof, "_first": first, "_nwin": nwin, "_chrom_per_rank": len(mine), "_nsteps": nsteps,
            "_lanczos_steps": lanczos_steps}


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_css(sample_windows, steps=1, warmup=0, mds=None):
    """the reference's own pthreads CSS (oracle/_ref/libref_css.so, 64 threads hard-wired) on a bounded sample:
    the first `sample_windows` windows of chromosome 0 of the headline workload"""
    import checkers
    import fpt_b200.synth as synth
    if not checkers.ref_available():
        return None
    ref = checkers.load_ref_css()
    ch = synth.chromosome_fast(CSS["seed0"], CSS["length"], CSS["nsnp"], CSS["asize"], CSS["bsize"], wstep=CSS["wstep"])
    regend = sample_windows * CSS["wstep"]
    keep = int(np.searchsorted(ch["pos"], regend + CSS["wsize"], side="right"))
    sub = {"pos": ch["pos"][:keep], "acodes": ch["acodes"][:keep * CSS["asize"]], "bcodes": ch["bcodes"][:keep * CSS["bsize"]],
           "asize": CSS["asize"], "bsize": CSS["bsize"]}
    av, bv, apos, bpos = synth.reference_layout(sub)
    n = regend // CSS["wstep"]
    times = []
    for it in range(warmup + steps):
        s, p = np.zeros(n + 8), np.zeros(n + 8)
        with checkers.silence_stdout():
            t0 = time.perf_counter()
            ref.threadcompute(checkers.dptr(av), checkers.dptr(bv), checkers.iptr(apos), checkers.iptr(bpos), 0, regend,
                              CSS["wsize"], CSS["wstep"], av.size, bv.size, CSS["mct"], CSS["mcr"], 0, CSS["mds"] if mds is None else mds,
                              checkers.dptr(s), checkers.dptr(p))
            dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return dict(seconds=times, windows=n, scores=s[:n])


def cpu_fet(sample_snps):
    import checkers
    import fpt_b200.synth as synth
    if not checkers.ref_available():
        return None
    ref = checkers.load_ref_fet()
    ch = synth.chromosome_fast(FET["seed0"], FET["length"], FET["nsnp"], FET["asize"], FET["bsize"], wstep=FET["wstep"])
    keep = min(sample_snps, ch["pos"].size)
    regend = int(ch["pos"][keep - 1]) // FET["wstep"] * FET["wstep"]
    sub = {"pos": ch["pos"][:keep], "acodes": ch["acodes"][:keep * FET["asize"]], "bcodes": ch["bcodes"][:keep * FET["bsize"]],
           "asize": FET["asize"], "bsize": FET["bsize"]}
    av, bv, apos, bpos = synth.reference_layout(sub)
    n = regend // FET["wstep"]
    s, d = np.zeros(n + 8), np.zeros(n + 8)
    with checkers.silence_stdout():
        t0 = time.perf_counter()
        ref.threadcompute(checkers.dptr(av), checkers.dptr(bv), checkers.iptr(apos), checkers.iptr(bpos), 0, regend, FET["wsize"],
                          FET["wstep"], av.size, bv.size, FET["perc"], checkers.dptr(s), checkers.dptr(d))
        dt = time.perf_counter() - t0
    nsnp_in = int(np.searchsorted(ch["pos"][:keep], regend, side="right"))
    return dict(seconds=dt, snps=nsnp_in, windows=n, scores=s[:n], regend=regend)


def cpu_fet_tables(n=200_000):
    """configs[3] lies outside the reference's u64 arithmetic, so the CPU figure is the oracle's log-space port,
    one core"""
    import checkers
    import fpt_b200.synth as synth
    o = checkers.load_oracle()
    T = synth.coverage_tables(FET_TABLES["seed0"], n, FET_TABLES["lo"], FET_TABLES["hi"])
    out = np.zeros(n)
    t0 = time.perf_counter()
    o.fpt_oracle_fet_tables(checkers.iptr(T), n, checkers.dptr(out))
    return n / (time.perf_counter() - t0), T, out


def cpu_large(first):
    """the reference's serial `compute` (css.c:49-156, oracle/_ref) on the first LARGE['cpu_windows'] windows of the large-cohort
    chromosome: one core (the pthreads driver hands out tasks of 100 windows, threadcss.c:58, so a sample this size would keep
    one thread busy anyway; 64 threads x 40 MB of matrices per thread is what makes it slow, not impossible)"""
    import checkers
    if not checkers.ref_available() or first is None:
        return None
    ref = checkers.load_ref_css()
    hpos, ha, hb, s_gpu, p_gpu, _ = first
    nw = LARGE["cpu_windows"]
    regend = nw * LARGE["wstep"]
    keep = int(np.searchsorted(hpos, regend + LARGE["wsize"], side="right"))
    asize, bsize = LARGE["asize"], LARGE["bsize"]

    def vals(c):
        v = c.astype(np.float64)
        v[c == -128] = -10000.0
        return v
    av, bv = vals(ha[:keep * asize]), vals(hb[:keep * bsize])
    apos, bpos = np.repeat(hpos[:keep], asize).astype(np.int32), np.repeat(hpos[:keep], bsize).astype(np.int32)
    s, p = np.zeros(nw + 8), np.zeros(nw + 8)
    with checkers.silence_stdout():
        t0 = time.perf_counter()
        ref.compute(checkers.dptr(av), checkers.dptr(bv), checkers.iptr(apos), checkers.iptr(bpos), 0, regend, LARGE["wsize"], LARGE["wstep"],
                    av.size, bv.size, LARGE["mcr"], LARGE["mcr"], 0, 0, checkers.dptr(s), checkers.dptr(p))
        dt = time.perf_counter() - t0
    both = (s[:nw] != 0) & (s_gpu[:nw] != 0)
    rel = np.abs(s[:nw][both] - s_gpu[:nw][both]) / np.maximum(np.abs(s[:nw][both]), 1e-300)
    return {"value": nw / dt, "unit": "windows/s", "cores": 1, "kind": "reference",
            "sample": "first %d windows of the large-cohort chromosome through the reference's serial compute() (%.1f s); css.c against the "
                      "header-only GSL stand-in" % (nw, dt),
            "parity_vs_gpu": {"windows_compared": int(both.sum()), "score_max_rel": float(rel.max()) if both.any() else None,
                              "score_rel_within_1e-5": float((rel <= 1e-5).mean()) if both.any() else None}}


def css_config(world):
    return {"workload": "CSS scan, ONE 450 Mb synthetic genome (BASELINE configs[2]): %d chromosomes x %d bp, %d SNPs each, "
                        "%d+%d individuals, wsize %d / wstep %d, classical MDS (mds=0), mcT=mcR=%d permutations; "
                        "genome inputs (1.44 GB float64) exceed L2, no flush" % (
                            CSS["chromosomes"], CSS["length"], CSS["nsnp"], CSS["asize"], CSS["bsize"], CSS["wsize"],
                            CSS["wstep"], CSS["mcr"]),
            "windows_per_step": CSS["chromosomes"] * (CSS["length"] // CSS["wstep"]),
            "sharding": "%d contiguous window range(s) of the one genome with SNP halos (sharding.py), one NCCL all_gather of scores + p "
                        "inside the timed region" % world}


def run_reference(args, rank, out):
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    probe = cpu_css(2000)                                 # calibrate the sample to a few minutes in total
    if probe is None:
        out.emit({"impl": "reference", "unavailable": "oracle/_ref/libref_css.so not built (reference tree absent at build time)"})
        return
    rate = probe["windows"] / probe["seconds"][0]
    budget = 150.0 / max(1, args.steps + args.warmup)
    nwin = int(max(2000, min(CSS["length"] // CSS["wstep"] - 8, rate * budget)))
    res = cpu_css(nwin, steps=args.steps, warmup=args.warmup)
    sec = float(np.mean(res["seconds"]))
    val = res["windows"] / sec
    sample = "first %d windows of chromosome 0 of the headline workload per step (bounded sample of the same config)" % res["windows"]
    line = {"impl": "reference", "metric": "css_windows_per_sec_1000perms", "value": val, "unit": "windows/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1000.0, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": css_config(max(1, args.gpus)),
            "cpu_baseline": {"value": val, "unit": "windows/s", "cores": min(64, ncores), "threads": 64, "nproc": ncores,
                             "kind": "reference", "sample": sample,
                             "note": "unmodified reference threadcompute (64 pthreads hard-wired), css.c linked against the "
                                     "header-only GSL stand-in of oracle/gsl_shim"},
            "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    out.emit(line)


class OneLineStdout:
    """everything any library prints to stdout while the benchmark runs (NCCL banners, reference printf()s) goes to
    stderr; the single JSON line is written to the real stdout at the end"""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, obj):
        sys.stdout.flush()
        os.write(self.real, (json.dumps(obj) + "\n").encode())


# ------------------------------------------------------------------------------------------------ rooflines
def frac(ach, peak):
    return (ach / peak) if (ach is not None and peak) else None


def css_kernel_table(css, micro, hbm, nsteps):
    """per kernel of the headline step (rank 0's share of it): time per launch, share, and algorithmic work against the measured
    peak of the resource that bounds it (DESIGN.md section 4 states each work model)"""
    m, a, b = CSS["asize"] + CSS["bsize"], CSS["asize"], CSS["bsize"]
    prof = css["prof"]
    tot = sum(v["ms"] for v in prof.values()) or 1.0
    nwin_rank, scored_rank = css["my_windows"], css["my_scored"]
    nsnp_rank = css["my_windows"] * CSS["wstep"] // 100       # 1 SNP / 100 bp
    fp64 = micro["fp64"]["tflops"] * 1e12 if micro else None
    smem = micro["smem"]["tb_per_s"] * 1e12 if micro else None
    issue = micro["issue"]["gwarp_inst_per_s"] * 1e9 if micro else None
    out = {}
    for k, v in prof.items():
        sec_step = v["ms"] * 1e-3 / nsteps                # this kernel's time per step (all its launches)
        e = {"ms_per_launch": v["ms"] / v["launches"], "share_of_step": round(v["ms"] / tot, 4)}
        if k == "css_perm":
            # R (asize bsize + m - 2) fp64 adds fed by as many 8-byte shared-memory gathers per scored window
            gath = scored_rank * CSS["mcr"] * (a * b + m - 2) * 8.0
            e.update(bound="smem", algorithmic_bytes_per_step=gath, achieved=gath / sec_step / 1e12, peak=smem and smem / 1e12, unit="TB/s",
                     frac=frac(gath / sec_step, smem),
                     note="work model of the reference algorithm (one 8-byte gather per score term); the kernel itself replaces most of "
                          "them by an exact integer surrogate on the u8 tensor cores, so frac > 1 is possible")
            ninst = ncu_record("css_perm", "inst")
            if ninst and issue and scored_rank:
                # the capture is the launch of chromosome 0 (NCU_CSS_SCORED scored windows): instructions per scored window x the
                # windows this rank scored per step, so that partial chromosomes (N > 1) and --small shapes stay comparable
                inst_step = ninst / NCU_CSS_SCORED * scored_rank
                e["issue"] = {"warp_inst_per_launch_ncu": ninst, "warp_inst_per_scored_window": ninst / NCU_CSS_SCORED,
                              "achieved_gwarp_inst_per_s": inst_step / sec_step / 1e9,
                              "peak": issue / 1e9, "frac": inst_step / sec_step / issue,
                              "note": "utilisation of the issue slots, from the committed ncu capture at N = 1 (not a roofline: useless instructions raise it)"}
        elif k in ("css_tridiag", "css_eigvec"):
            # K5: double centring ~6 m^2 + dense symmetric eigen-decomposition ~9 m^3 (what the reference's GSL call does); split 2/3 : 1/3
            fl = scored_rank * ((6.0 * m * m + 9.0 * m ** 3) * (2.0 / 3.0 if k == "css_tridiag" else 1.0 / 3.0))
            e.update(bound="fp64", algorithmic_flops_per_step=fl, achieved=fl / sec_step / 1e12, peak=fp64 and fp64 / 1e12, unit="TFLOP/s",
                     frac=frac(fl / sec_step, fp64),
                     note="dense work model 6 m^2 + 9 m^3 per window (reduction 2/3, eigenvectors 1/3); the kernels do less than that "
                          "(two eigenpairs only) and are latency / issue bound at one warp per 40 x 40 window")
        elif k == "css_observed":
            # the observed score: asize bsize + m - 2 DEPENDENT fp64 additions per scored window (css.c:608-647), one warp per window
            adds = scored_rank * float(a * b + m - 2)
            e.update(bound="fp64 add latency", algorithmic_flops_per_step=adds,
                     note="a chain of dependent additions in the reference's order: the bound is the add latency times the chain length "
                          "over the windows resident at once, not a throughput peak")
        elif k == "css_pack":
            by = nsnp_rank * (m * 8 + m / 4.0)
            e.update(bound="hbm", algorithmic_bytes_per_step=by, achieved=by / sec_step / 1e9, peak=hbm, unit="GB/s", frac=frac(by / sec_step / 1e9, hbm),
                     traffic=ncu_record("css_pack"))
        elif k == "window_table":
            by = nwin_rank * 8.0
            e.update(bound="latency", algorithmic_bytes_per_step=by, note="two binary searches per window over L2-resident positions")
        out[k] = e
    return out


def large_kernel_table(large, micro, hbm):
    """per kernel of the large-cohort leg (rank 0's chromosomes): time per chromosome, share, and the work models of the Lanczos and
    permutation kernels. The code route runs a chromosome in passes of <= 4096 windows (one profile scope per pass; 1024 in earlier builds), so every
    model is set against the kernel's time per CHROMOSOME (all its passes), never per scope."""
    prof = large.pop("_prof")
    nwin, cpr, nsteps = large.pop("_nwin"), large.pop("_chrom_per_rank"), large.pop("_nsteps")
    lz_steps = large.pop("_lanczos_steps")
    mL = LARGE["asize"] + LARGE["bsize"]
    ltot = sum(v["ms"] for v in prof.values()) or 1.0
    passes = max(1, nsteps * cpr)                         # chromosomes this rank ran while the profile was on
    lk = {}
    for k, v in prof.items():
        per = v["ms"] / passes
        e = {"ms_per_chromosome": per, "ms_per_launch": v["ms"] / v["launches"], "scopes_per_chromosome": v["launches"] / passes,
             "share_of_step": round(v["ms"] / ltot, 4)}
        if k == "css_mds_large":
            # algorithmic HBM bytes per window: the window's bit-planes in (2 planes x m x ceil(npos/32) words) and the embedding out (16 m)
            words = -(-LARGE["snps_per_window"] // 32) + 1
            by = nwin * (2.0 * mL * words * 4 + 16.0 * mL)
            # the committed ncu capture (profiles/capture.sh) is ONE launch over NCU_LARGE_WINDOWS windows: scale it to a chromosome
            tr = ncu_record("css_mds_large")
            tr = tr * nwin / NCU_LARGE_WINDOWS if tr else None
            e.update(bound="hbm", algorithmic_bytes=by, achieved=by / (per * 1e-3) / 1e9, peak=hbm, unit="GB/s", frac=by / (per * 1e-3) / 1e9 / hbm,
                     traffic=tr, traffic_over_algorithmic=(tr / by) if tr else None,
                     note="per chromosome of %d windows: 72 KB of input + output per window; everything above that in `traffic` (ncu capture "
                          "of %d windows, scaled by the window count) is the kernel re-streaming its own code matrix and Lanczos basis"
                          % (nwin, NCU_LARGE_WINDOWS))
            if micro and lz_steps:
                # what bounds it: issue slots and the fp64 pipe of the product. Work of the Krylov method itself per window: steps x
                # (2 m^2 for the product + 4 m x (mean basis size = steps / 2) for the Gram-Schmidt pass); the dense 9 m^3 of the
                # reference's solver is not what this kernel does, so it is not used as a roofline
                fl = nwin * lz_steps * (2.0 * mL * mL + 4.0 * mL * lz_steps / 2.0)
                e.update(bound="fp64", lanczos_steps_per_window=lz_steps,
                         fp64={"algorithmic_flops": fl, "achieved": fl / (per * 1e-3) / 1e12, "peak": micro["fp64"]["tflops"],
                               "frac": fl / (per * 1e-3) / 1e12 / micro["fp64"]["tflops"], "unit": "TFLOP/s",
                               "note": "Lanczos work model: steps x (2 m^2 + 2 m steps) flops per window, steps counted by the kernel; the "
                                       "product spends a second fp64 instruction per element on turning c^2 into a double"},
                         hbm={"algorithmic_bytes": by, "frac": by / (per * 1e-3) / 1e9 / hbm, "traffic": tr,
                              "traffic_over_algorithmic": (tr / by) if tr else None})
        elif k == "css_perm" and micro:
            npad, kpad, batches = -(-mL // 256) * 256, -(-mL // 128) * 128, -(-LARGE["mcr"] // 128)
            macs = float(nwin) * batches * 128 * npad * kpad * 4
            e.update(bound="tensor", algorithmic_macs=macs, achieved=2.0 * macs / (per * 1e-3) / 1e12, peak=micro["umma_i8"]["tops"], unit="TOP/s",
                     frac=2.0 * macs / (per * 1e-3) / 1e12 / micro["umma_i8"]["tops"],
                     note="u8 contraction work of the kernel (4 base-256 digits) over the WHOLE kernel time, against the kind::i8 rate measured "
                          "with both operands resident in shared memory (profiles/microbench/peaks.cu)")
        lk[k] = e
    large["kernels"] = lk
    large["kernel_ms_per_chromosome"] = {k: v["ms"] / passes for k, v in prof.items()}
    return large


def main():
    out = OneLineStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-fet", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-large", action="store_true", help="skip the 500+500 cohort chromosomes")
    ap.add_argument("--skip-replicas", action="store_true", help="N > 1: skip the secondary one-genome-per-GPU figure")
    ap.add_argument("--small", action="store_true", help="tiny shapes: checks that the script runs, not a benchmark")
    ap.add_argument("--chromosomes", type=int, default=None, help="profiling aid: fewer chromosomes than the 21 of the workload")
    args = ap.parse_args()
    if args.small:
        CSS.update(chromosomes=2, length=2_000_000, nsnp=20_000)
        FET.update(length=10_000_000, nsnp=100_000)
        FET_TABLES.update(n=1_600_000)
    if args.chromosomes:
        CSS.update(chromosomes=args.chromosomes)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, out)
        return
    import torch
    import fpt_b200._lib as lib_mod
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = lib_mod.load()
    lib.fpt_set_seed(SEED)
    lib.fpt_set_device(local)
    micro = micro_peaks() if rank == 0 else None
    cx = Ctx(args, rank, world, dist)
    css = bench_css(lib_mod, cx, full_genome_pass=(world > 1 and not args.skip_replicas))
    fet = fet_tab = None
    if not args.skip_fet:
        fet, fet_sample = bench_fet(lib_mod, cx)
        fet_tab = bench_fet_tables(lib_mod, cx)
    large = None
    if not args.skip_large:
        large = bench_large_cohort(lib_mod, cx)
    if world > 1:
        dist.barrier()
    if rank == 0:
        hbm, hbm_src = hbm_peak()
        m = CSS["asize"] + CSS["bsize"]
        ms_step = css["ms"] / args.steps
        value = css["total_windows"] / (ms_step * 1e-3)
        e2e_val = css["total_windows"] / (css["ms_e2e"] / args.steps * 1e-3)
        kernels = css_kernel_table(css, micro, hbm, args.steps)
        dom = max(css["prof"], key=lambda k: css["prof"][k]["ms"])
        kd = kernels[dom]
        line = {
            "metric": "css_windows_per_sec_1000perms", "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": css_config(world),
            "windows_scored_per_step": css["scored"],
            "perm_rechecks": {"exact_rescorings_rank0": css["rechecks"],
                              "note": "permutations whose integer surrogate score was too close to the observed score to decide "
                                      "`>=` and were re-scored in the reference's summation order (rank 0, all steps so far)"},
            "e2e": {"value": e2e_val, "unit": "windows/s", "h2d_bytes_per_step": css["h2d"], "d2h_bytes_per_step": css["d2h"],
                    "api": ("fpt_css_compute (drop-in, host float64 arrays in pinned memory), one call per chromosome" if world == 1 else
                            "fpt_css_scan on every rank's window range (host float64 arrays in pinned memory) + one NCCL all_gather + D2H of the genome's results"),
                    "windows_scored": css["scored_e2e"]},
            "e2e_pageable": css["pageable"],
            "gpu_launches": css["launches"],
            "roofline": {"kernel": dom, "bound": kd.get("bound"), "achieved": kd.get("achieved"), "peak": kd.get("peak"), "unit": kd.get("unit"),
                         "frac": kd.get("frac"), "traffic": ncu_record(dom),
                         "peak_source": (micro or {}).get("source", "unavailable"),
                         "share_of_step": {k: v["share_of_step"] for k, v in kernels.items()},
                         "kernel_ms_per_launch": {k: v["ms_per_launch"] for k, v in kernels.items()},
                         "note": kd.get("note")},
            "kernels": kernels,
            "peaks": {"hbm_gbs": hbm, "hbm_source": hbm_src, "micro": micro},
            "clocks": css["clocks"],
            "mds_variants": {"note": "same workload, chromosome 0 only, device-resident (rank 0): mds=1 is SMACOF from 4 random starts, mds=2 is "
                                     "classical MDS followed by SMACOF (css.c:208-218)", **css["variants"]},
        }
        # K6 roofline of the SMACOF kernel: iterations x 14 m^2 fp64 flops
        if micro:
            for kname, v in css["variants"].items():
                ms_k = v["kernel_ms"].get("css_smacof")
                if ms_k:
                    fl = v["smacof_iterations_total"] * 14.0 * m * m
                    v["roofline_smacof"] = {"bound": "fp64", "algorithmic_flops": fl, "achieved": fl / (ms_k * 1e-3) / 1e12,
                                            "peak": micro["fp64"]["tflops"], "unit": "TFLOP/s", "frac": fl / (ms_k * 1e-3) / 1e12 / micro["fp64"]["tflops"]}
        if css["replicas"]:
            line["replicas"] = css["replicas"]
        if world == 1 and not args.skip_cpu:
            cb = cpu_css(1000)
            if cb is not None:                           # scale the sample to roughly 15 s of CPU wall time
                want = int(min(CSS["length"] // CSS["wstep"] - 8, max(1000, 15.0 * cb["windows"] / cb["seconds"][0])))
                if want > 1500:
                    cb = cpu_css(want)
            if cb is not None:
                sec = cb["seconds"][0]
                ncores = os.cpu_count() or 1
                ch0, host0, s0, p0 = css["sample"]
                ref_s = cb["scores"]
                n = cb["windows"]
                both = (ref_s != 0) & (s0[:n] != 0) & np.isfinite(ref_s)
                rel = np.abs(s0[:n][both] - ref_s[both]) / np.maximum(np.abs(ref_s[both]), 1e-300)
                line["cpu_baseline"] = {"value": n / sec, "unit": "windows/s", "cores": min(64, ncores), "threads": 64, "nproc": ncores,
                                        "kind": "reference",
                                        "sample": "first %d windows of chromosome 0 of the headline workload (%.1f s of CPU wall time)" % (n, sec),
                                        "parity_vs_gpu": {"windows_compared": int(both.sum()),
                                                          "same_windows_scored": bool(np.array_equal(ref_s[:n - 8] != 0, s0[:n - 8] != 0)),
                                                          "score_rel_within_1e-5": float((rel <= 1e-5).mean()) if both.any() else None,
                                                          "score_max_rel": float(rel.max()) if both.any() else None}}
                # the other two MDS variants: the reference's threadcompute with mds = 1 / 2 on a smaller sample (SMACOF is ~10-40x the work);
                # mds = 2 starts from classical MDS, so its scores are comparable window by window; mds = 1 starts from clock-seeded
                # random configurations in the reference (css.c:863-864), so only its time is reported
                for mds_v, nsamp in ((2, 20000), (1, 6000)):
                    v = css["variants"].get("mds%d" % mds_v)
                    if not v:
                        continue
                    cv = cpu_css(nsamp, mds=mds_v)
                    if not cv:
                        continue
                    secv = cv["seconds"][0]
                    e = {"value": cv["windows"] / secv, "unit": "windows/s", "cores": min(64, ncores), "threads": 64, "kind": "reference",
                         "sample": "first %d windows of chromosome 0, reference threadcompute with mds = %d (%.1f s)" % (cv["windows"], mds_v, secv)}
                    gs = v.get("_scores")
                    if mds_v == 2 and gs is not None:
                        nn = cv["windows"]
                        rs = cv["scores"]
                        bothv = (rs != 0) & (gs[:nn] != 0) & np.isfinite(rs)
                        relv = np.abs(gs[:nn][bothv] - rs[bothv]) / np.maximum(np.abs(rs[bothv]), 1e-300)
                        e["parity_vs_gpu"] = {"windows_compared": int(bothv.sum()), "score_rel_within_1e-5": float((relv <= 1e-5).mean()) if bothv.any() else None,
                                              "score_max_rel": float(relv.max()) if bothv.any() else None}
                    v["cpu_baseline"] = e
            else:
                line["cpu_baseline"] = {"value": None, "unit": "windows/s", "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref not available"}
        for v in css["variants"].values():
            v.pop("_scores", None)
        if fet is not None:
            prof = fet.pop("_prof")
            ns, nw = fet.pop("_ns"), fet.pop("_nw")
            asz = FET["asize"] + FET["bsize"]
            ftot = sum(v["ms"] for v in prof.values()) or 1.0
            kb = {"fet_count": ns * (asz * 8 + 16), "fet_score": ns * 24, "window_table": nw * 8}
            fk = {}
            for k, v in prof.items():
                per = v["ms"] / v["launches"]
                e = {"ms_per_launch": per, "share_of_step": round(v["ms"] / ftot, 4)}
                if k in kb and k != "window_table":
                    e.update(bound="hbm", algorithmic_bytes=kb[k], achieved=kb[k] / (per * 1e-3) / 1e9, peak=hbm, unit="GB/s",
                             frac=kb[k] / (per * 1e-3) / 1e9 / hbm, traffic=ncu_record(k))
                elif k == "fet_window" and micro:
                    # K3: 100 x npos draws per window, each one counter read-modify-write (2 x 2 B) plus the counting pass; report
                    # draws/s and the shared-memory bytes they imply against the measured shared-memory peak
                    draws = 100.0 * ns * (FET["wsize"] // FET["wstep"])
                    e.update(bound="smem/issue", bootstrap_draws=draws, draws_per_s=draws / (per * 1e-3),
                             smem_bytes_min=draws * 4, achieved=draws * 4 / (per * 1e-3) / 1e12, peak=micro["smem"]["tb_per_s"], unit="TB/s",
                             frac=draws * 4 / (per * 1e-3) / 1e12 / micro["smem"]["tb_per_s"],
                             note="integer / shared-memory bound (48-bit LCG draw + counter update per bootstrap index); no meaningful HBM fraction")
                fk[k] = e
            fet["kernels"] = fk
            domf = max(prof, key=lambda k: prof[k]["ms"])
            fet["roofline"] = dict(kernel=domf, **{k: v for k, v in fk[domf].items() if k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "note")})
            if world == 1 and not args.skip_cpu:
                cf = cpu_fet(200_000)
                if cf is not None:
                    ncores = os.cpu_count() or 1
                    ch, s_host, wb, we = fet_sample
                    n = cf["windows"]
                    fet["cpu_baseline"] = {"value": cf["snps"] / cf["seconds"], "unit": "SNPs/s", "cores": min(64, ncores), "threads": 64,
                                           "nproc": ncores, "kind": "reference",
                                           "sample": "first %d SNPs (%d windows) of the FET workload, reference threadcompute, %.1f s" % (
                                               cf["snps"], n, cf["seconds"]),
                                           "parity_vs_gpu": {"score_max_abs_diff": float(np.max(np.abs(s_host[:n - 8] - cf["scores"][:n - 8])))}}
            line["fet"] = fet
        if fet_tab is not None:
            k_ms, work, n_loc = fet_tab.pop("_kernel_ms"), fet_tab.pop("_work"), fet_tab.pop("_n_local_max")
            rf = {"kernel": "fet_score (log mode)", "ms_per_launch": k_ms, "tables_per_launch": n_loc}
            if k_ms:
                rf["hbm"] = {"bound": "hbm", "algorithmic_bytes": n_loc * 24, "achieved": n_loc * 24 / (k_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                             "frac": n_loc * 24 / (k_ms * 1e-3) / 1e9 / hbm}
                if work and micro:
                    # 6 T + 12 flops per table (and T divides, counted as one flop each here)
                    fl = (7.0 * work["mean_T"] + 12.0) * n_loc
                    rf["fp64"] = {"bound": "fp64", "algorithmic_flops": fl, "achieved": fl / (k_ms * 1e-3) / 1e12, "peak": micro["fp64"]["tflops"],
                                  "unit": "TFLOP/s", "frac": fl / (k_ms * 1e-3) / 1e12 / micro["fp64"]["tflops"],
                                  "note": "work of the reference's full two-tailed walk (7 T + 12 per table, divides counted as one); the kernel skips "
                                          "tail terms below 2^-60 by bisection, so it does less"}
                rf["bound"] = "fp64" if ("fp64" in rf and rf["fp64"]["frac"] > rf["hbm"]["frac"]) else "hbm"
            fet_tab["roofline"] = rf
            fet_tab["work_model"] = work
            if world == 1 and not args.skip_cpu:
                rate, Tc, oc = cpu_fet_tables()
                g = np.zeros(len(Tc))
                from fpt_b200._lib import check
                check(lib.fpt_fet_tables(Tc.ctypes.data, len(Tc), 0, g.ctypes.data))
                rel = np.abs(g - oc) / np.maximum(np.abs(oc), 1e-300)
                fet_tab["cpu_baseline"] = {"value": rate, "unit": "SNPs/s", "cores": 1, "kind": "port",
                                           "sample": "%d tables of the same distribution through the oracle's log-space restatement (the reference's "
                                                     "u64 binomials overflow beyond N = 67, SURVEY Q2)" % len(Tc),
                                           "parity_vs_gpu": {"max_rel_where_score_above_1e-6": float(rel[np.abs(oc) > 1e-6].max()),
                                                             "max_abs": float(np.abs(g - oc).max()),
                                                             "within_1e-9_rel_plus_1e-12_abs": bool(np.all(np.abs(g - oc) <= 1e-9 * np.abs(oc) + 1e-12))}}
            line["fet_tables"] = fet_tab
        if large is not None:
            first = large.pop("_first")
            large_kernel_table(large, micro, hbm)
            if world == 1 and not args.skip_cpu and not args.small:
                large["cpu_baseline"] = cpu_large(first)
            line["large_cohort"] = large
        out.emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
