#!/usr/bin/env python
"""bench.py — throughput of the two divergence scans on B200, with the reference's CPU path timed beside it.

Headline workload (BASELINE.json configs[2], SURVEY.md 8(d) "C3"): CSS over a 450 Mb stickleback-sized synthetic
genome — 21 chromosomes, 1 SNP / 100 bp (4.5 M SNPs), 20+20 diploid individuals, wsize 2500 / wstep 500
(~900 k windows), classical MDS, mcT = mcR = 1000 permutations per window. One step = one pass over the genome,
one scan call per chromosome (the reference's call granularity).

  value  windows/s with the float64 genotype arrays already resident in HBM (device API, CUDA events)
  e2e    windows/s through the drop-in host call fpt_css_compute (host -> device copies of every input from pinned
         host memory and device -> host reads of the results inside the timed region)
  fet    the same pair for the FET scan (BASELINE configs[0] geometry, 1 M SNPs, 20+20, 2500/500, perc 0.95)
         plus direct 2x2 tables with coverage <= 500 (configs[3] per-GPU shard, 12.5 M tables)

`--impl reference` times the reference's own pthreads C (oracle/_ref, compiled unmodified from the reference tree)
on the host cores, on a bounded sample of the same workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
import ctypes as C
import json
import os
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
FET_TABLES = dict(n=12_500_000, lo=20, hi=500, seed0=20261018 + 3)
SEED = 20261018


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, what="traffic"):
    """per-launch DRAM bytes (or warp instructions, what="inst") of `kernel` from the committed ncu capture, if one has
    been summarised"""
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


# ------------------------------------------------------------------------------------------------ data
def make_css_genome(rank):
    import fpt_b200.synth as synth
    chroms = []
    for c in range(CSS["chromosomes"]):
        chroms.append(synth.chromosome_fast(CSS["seed0"] + 1000 * rank + c, CSS["length"], CSS["nsnp"], CSS["asize"],
                                            CSS["bsize"], wstep=CSS["wstep"]))
    return chroms


def pinned_like(arr):
    import torch
    t = torch.empty(arr.shape, dtype=torch.from_numpy(arr[:0]).dtype).pin_memory()
    v = t.numpy()
    v[...] = arr
    return t, v


# ------------------------------------------------------------------------------------------------ B200 arm: CSS
def bench_css(lib_mod, args, rank, world, dist):
    import torch
    import fpt_b200.synth as synth
    from fpt_b200._lib import ScanRange, check
    lib = lib_mod.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    chroms = make_css_genome(rank)
    asize, bsize, m = CSS["asize"], CSS["bsize"], CSS["asize"] + CSS["bsize"]
    regend, wsize, wstep = CSS["length"], CSS["wsize"], CSS["wstep"]
    nout = regend // wstep
    host, resident = [], []
    for ch in chroms:                                    # reference layout: float64 values + repeated int32 positions
        av, bv, apos, bpos = synth.reference_layout(ch)
        hp = [pinned_like(x) for x in (av, bv, apos, bpos)]
        host.append(hp)
        resident.append((hp[0][0].to(dev), hp[1][0].to(dev), torch.from_numpy(ch["pos"]).to(dev)))
    nsnp = CSS["nsnp"]
    planes = torch.empty(lib.fpt_dev_css_planes_bytes(nsnp, m) // 4, dtype=torch.int32, device=dev)
    wl = torch.empty(nout, dtype=torch.int32, device=dev)
    wr = torch.empty(nout, dtype=torch.int32, device=dev)
    mx = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.fpt_dev_css_workspace_bytes(m, nout, CSS["mds"])
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    out_s = [torch.zeros(nout, dtype=torch.float64, device=dev) for _ in chroms]
    out_p = [torch.zeros(nout, dtype=torch.float64, device=dev) for _ in chroms]
    out_st = [torch.zeros(nout, dtype=torch.uint8, device=dev) for _ in chroms]
    gathered = None
    if world > 1:
        gathered = torch.empty(world * len(chroms) * nout * 2, dtype=torch.float64, device=dev)
    r = ScanRange()
    r.regend, r.wsize, r.wstep, r.semantics, r.window_begin, r.window_end, r.seed = regend, wsize, wstep, 0, 0, nout, SEED
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def step_resident():
        for i, (da, db, dpos) in enumerate(resident):
            check(lib.fpt_dev_css_pack_f64(da.data_ptr(), db.data_ptr(), nsnp, asize, bsize, planes.data_ptr(), sp))
            mx.zero_()
            check(lib.fpt_dev_window_table(dpos.data_ptr(), nsnp, C.byref(r), wl.data_ptr(), wr.data_ptr(), mx.data_ptr(), sp))
            out_st[i].zero_()
            check(lib.fpt_dev_css_windows(planes.data_ptr(), None, asize, bsize, wl.data_ptr(), wr.data_ptr(), C.byref(r),
                                          CSS["mct"], CSS["mcr"], CSS["mds"], ws.data_ptr(), ws_bytes, out_s[i].data_ptr(),
                                          out_p[i].data_ptr(), out_st[i].data_ptr(), None, sp))
        if world > 1:                                    # the only exchange: results gathered once at the end
            local = torch.cat([torch.cat(out_s), torch.cat(out_p)])
            dist.all_gather_into_tensor(gathered, local)

    launches_per_step = len(chroms) * 5                  # pack, window table, tridiagonalisation, eigenvectors, permutations

    def step_e2e():
        total = 0
        for (av, bv, apos, bpos) in host:
            s = np.zeros(nout)
            p = np.zeros(nout)
            check(lib.fpt_css_compute(av[1].ctypes.data, bv[1].ctypes.data, apos[1].ctypes.data, bpos[1].ctypes.data, 0, regend,
                                      wsize, wstep, av[1].size, bv[1].size, CSS["mct"], CSS["mcr"], 0, CSS["mds"],
                                      s.ctypes.data, p.ctypes.data))
            total += int(np.count_nonzero(p))
        return total

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    lib.fpt_profile_enable(1)
    buf = C.create_string_buffer(4096)
    lib.fpt_profile_summary(buf, 4096)                   # clear
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(torch.cuda.current_device()) as clk:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_resident()
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    lib.fpt_profile_summary(buf, 4096)
    lib.fpt_profile_enable(0)
    prof = json.loads(buf.value.decode())
    scored = int(sum(int((st == 2).sum().item()) for st in out_st))
    rechecks = int(lib.fpt_css_perm_rechecks())             # exact re-scorings since the library was loaded (all steps so far)
    windows_per_step = len(chroms) * nout                # window slots visited per rank and step

    # the other two MDS variants of BASELINE configs[2], on chromosome 0 only (device-resident, one warm-up + one timed pass each)
    variants = {}
    if not args.small or True:
        da, db, dpos = resident[0]
        check(lib.fpt_dev_css_pack_f64(da.data_ptr(), db.data_ptr(), nsnp, asize, bsize, planes.data_ptr(), sp))
        mx.zero_()
        check(lib.fpt_dev_window_table(dpos.data_ptr(), nsnp, C.byref(r), wl.data_ptr(), wr.data_ptr(), mx.data_ptr(), sp))
        for mds_v in (1, 2):
            wsb = lib.fpt_dev_css_workspace_bytes(m, nout, mds_v)
            wsv = torch.empty(wsb, dtype=torch.uint8, device=dev)
            sv, pv, stv = (torch.zeros(nout, dtype=torch.float64, device=dev), torch.zeros(nout, dtype=torch.float64, device=dev),
                           torch.zeros(nout, dtype=torch.uint8, device=dev))
            for it in range(2):
                stv.zero_()
                torch.cuda.synchronize()
                e0.record(stream)
                check(lib.fpt_dev_css_windows(planes.data_ptr(), None, asize, bsize, wl.data_ptr(), wr.data_ptr(), C.byref(r),
                                              CSS["mct"], CSS["mcr"], mds_v, wsv.data_ptr(), wsb, sv.data_ptr(), pv.data_ptr(),
                                              stv.data_ptr(), None, sp))
                e1.record(stream)
                torch.cuda.synchronize()
            variants["mds%d" % mds_v] = {"windows_per_s": nout / (e0.elapsed_time(e1) * 1e-3), "ms_per_chromosome": e0.elapsed_time(e1),
                                         "windows_scored": int((stv == 2).sum().item())}
            del wsv
    lib.fpt_profile_summary(buf, 4096)                   # drop the variants' launches from the per-kernel record

    # end to end through the drop-in host call
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        scored_e2e = step_e2e()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    tt = torch.tensor([ms, t_e2e * 1000.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(tt[0].item()), float(tt[1].item())
    h2d = sum(av[1].nbytes + bv[1].nbytes + 4 * nsnp for (av, bv, apos, bpos) in host)
    d2h = len(chroms) * nout * 17
    return dict(ms=ms, ms_e2e=ms_e2e, windows_per_step=windows_per_step, scored=scored, scored_e2e=scored_e2e, prof=prof,
                launches=launches_per_step * args.steps, h2d=h2d, d2h=d2h, rechecks=rechecks, variants=variants, clocks=clk.summary(),
                sample=(chroms[0], host[0], out_s[0].cpu().numpy(), out_p[0].cpu().numpy()))


def css_kernel_bytes(name, windows, nsnp_total, m):
    """algorithmic HBM bytes of one step, per kernel (DESIGN.md section 'Kernels and rooflines')"""
    if name == "css_perm":      # embedding in (16 m), status in (1), score + p out (16)
        return windows * (16 * m + 17)
    if name == "css_tridiag":   # the window's bit-plane slab in, window bounds (8); tridiagonal (24 m) + reflectors (4 m (m-1)) + status out
        return windows * (2 * 2 * m * 4 + 8 + 24 * m + 4 * m * (m - 1) + 1)
    if name == "css_eigvec":    # tridiagonal + reflectors + status in, embedding (16 m) out
        return windows * (24 * m + 4 * m * (m - 1) + 1 + 16 * m)
    if name == "css_mds_large":  # Lanczos kernel, ~80 steps: the matrix as 8-bit codes per step + four passes over the growing basis
        return windows * (80 * m * m + 4 * 8 * m * (80 * 81 // 2))
    if name == "css_pack":      # float64 genotypes in, two bit-planes out
        return nsnp_total * m * 8 + nsnp_total * m // 4
    if name == "window_table":  # positions are binary-searched (L2 resident); two int32 out
        return windows * 8
    return 0


# ------------------------------------------------------------------------------------------------ B200 arm: FET
def bench_fet(lib_mod, args):
    import torch
    import fpt_b200.synth as synth
    from fpt_b200._lib import ScanRange, check
    lib = lib_mod.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    ch = synth.chromosome_fast(FET["seed0"], FET["length"], FET["nsnp"], FET["asize"], FET["bsize"], wstep=FET["wstep"])
    av, bv, apos, bpos = synth.reference_layout(ch)
    hp = [pinned_like(x) for x in (av, bv, apos, bpos)]
    del av, bv, apos, bpos
    nsnp, asize, bsize = FET["nsnp"], FET["asize"], FET["bsize"]
    regend, wsize, wstep = FET["length"], FET["wsize"], FET["wstep"]
    nout = regend // wstep
    da, db = hp[0][0].to(dev), hp[1][0].to(dev)
    dpos = torch.from_numpy(ch["pos"]).to(dev)
    tab = torch.empty(nsnp * 4, dtype=torch.int32, device=dev)
    snp = torch.empty(nsnp, dtype=torch.float64, device=dev)
    wl = torch.empty(nout, dtype=torch.int32, device=dev)
    wr = torch.empty(nout, dtype=torch.int32, device=dev)
    mx = torch.zeros(1, dtype=torch.int32, device=dev)
    osc = torch.zeros(nout, dtype=torch.float64, device=dev)
    osd = torch.zeros(nout, dtype=torch.float64, device=dev)
    ofl = torch.zeros(nout, dtype=torch.uint8, device=dev)
    r = ScanRange()
    r.regend, r.wsize, r.wstep, r.semantics, r.window_begin, r.window_end, r.seed = regend, wsize, wstep, 0, 0, nout, SEED
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > L2 (126 MB)

    def step():
        check(lib.fpt_dev_fet_count_f64(da.data_ptr(), db.data_ptr(), nsnp, asize, bsize, tab.data_ptr(), sp))
        check(lib.fpt_dev_fet_score(tab.data_ptr(), nsnp, asize + bsize, 0, snp.data_ptr(), sp))
        mx.zero_()
        check(lib.fpt_dev_window_table(dpos.data_ptr(), nsnp, C.byref(r), wl.data_ptr(), wr.data_ptr(), mx.data_ptr(), sp))
        max_npos = int(mx.item())                        # the one host read the scan needs (sizes shared memory)
        ofl.zero_()
        check(lib.fpt_dev_fet_windows(snp.data_ptr(), wl.data_ptr(), wr.data_ptr(), C.byref(r), max_npos, FET["perc"], None,
                                      osc.data_ptr(), osd.data_ptr(), ofl.data_ptr(), sp))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    lib.fpt_profile_enable(1)
    buf = C.create_string_buffer(4096)
    lib.fpt_profile_summary(buf, 4096)
    tot = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.steps):
        flush.zero_()                                    # L2 flush between timed iterations (inputs are 320 MB anyway)
        torch.cuda.synchronize()
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    lib.fpt_profile_summary(buf, 4096)
    lib.fpt_profile_enable(0)
    prof = json.loads(buf.value.decode())
    ms = tot / args.steps

    def e2e():
        s, d = np.zeros(nout), np.zeros(nout)
        check(lib.fpt_fet_compute(hp[0][1].ctypes.data, hp[1][1].ctypes.data, hp[2][1].ctypes.data, hp[3][1].ctypes.data, 0,
                                  regend, wsize, wstep, hp[0][1].size, hp[1][1].size, FET["perc"], s.ctypes.data, d.ctypes.data))
        return s
    for _ in range(2):
        e2e()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s_host = e2e()
    ms_e2e = (time.perf_counter() - t0) * 1000.0 / args.steps
    hbm, hbm_src = peaks()
    kbytes = {"fet_count": nsnp * ((asize + bsize) * 8 + 16), "fet_score": nsnp * 24,
              "fet_window": nout * 24 + nsnp * 8 * (wsize // wstep), "window_table": nout * 8}
    dom = max(prof, key=lambda k: prof[k]["ms"])
    share = {k: round(v["ms"] / sum(x["ms"] for x in prof.values()), 4) for k, v in prof.items()}
    per_launch_ms = prof[dom]["ms"] / prof[dom]["launches"]
    ach = kbytes.get(dom, 0) / (per_launch_ms * 1e-3) / 1e9
    cnt_ms = prof["fet_count"]["ms"] / prof["fet_count"]["launches"]
    out = {
        "metric": "fet_snps_per_sec", "unit": "SNPs/s", "value": nsnp / (ms * 1e-3), "ms_per_step": ms,
        "config": {"workload": "FET scan, 1 chromosome 100 Mb, 1M SNPs, 20+20, wsize 2500 / wstep 500, perc 0.95 "
                               "(BASELINE configs[0] geometry); L2 flushed between steps", "windows": nout},
        "dtype": "f64",
        "e2e": {"value": nsnp / (ms_e2e * 1e-3), "unit": "SNPs/s", "h2d_bytes_per_step": hp[0][1].nbytes + hp[1][1].nbytes + 4 * nsnp,
                "d2h_bytes_per_step": nout * 17, "api": "fpt_fet_compute (drop-in, host float64 arrays)"},
        "gpu_launches": 4 * args.steps,
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                     "peak_source": hbm_src, "traffic": ncu_traffic(dom), "share_of_step": share,
                     "note": "fet_window is integer/shared-memory bound (100 x npos LCG draws per window); "
                             "the HBM-bound stage is fet_count, listed in roofline_hbm_stage"},
        "roofline_hbm_stage": {"kernel": "fet_count", "bound": "hbm", "achieved": kbytes["fet_count"] / (cnt_ms * 1e-3) / 1e9,
                               "peak": hbm, "unit": "GB/s", "frac": kbytes["fet_count"] / (cnt_ms * 1e-3) / 1e9 / hbm,
                               "traffic": ncu_traffic("fet_count")},
    }
    # direct 2x2 tables, coverage <= 500 (BASELINE configs[3], one GPU's shard of 100 M)
    n = FET_TABLES["n"]
    g = torch.Generator(device=dev)
    g.manual_seed(FET_TABLES["seed0"])
    n1 = torch.randint(FET_TABLES["lo"], FET_TABLES["hi"] + 1, (n,), device=dev, generator=g).double()
    n2 = torch.randint(FET_TABLES["lo"], FET_TABLES["hi"] + 1, (n,), device=dev, generator=g).double()
    f = torch.rand(n, device=dev, generator=g, dtype=torch.float64).clamp(0.02, 0.98)
    fb = (f + 0.05 * torch.randn(n, device=dev, generator=g, dtype=torch.float64)).clamp(0.01, 0.99)
    a = torch.binomial(n1, f, generator=g)
    c = torch.binomial(n2, fb, generator=g)
    T = torch.stack([a, n1 - a, c, n2 - c], dim=1).to(torch.int32).contiguous()
    del n1, n2, f, fb, a, c
    outp = torch.empty(n, dtype=torch.float64, device=dev)
    for _ in range(args.warmup):
        check(lib.fpt_dev_fet_score(T.data_ptr(), n, 1000, 0, outp.data_ptr(), sp))
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0.record(stream)
        check(lib.fpt_dev_fet_score(T.data_ptr(), n, 1000, 0, outp.data_ptr(), sp))
        e1.record(stream)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms_t = tot / args.steps
    out["tables"] = {"metric": "fet_snps_per_sec", "value": n / (ms_t * 1e-3), "unit": "SNPs/s", "ms_per_step": ms_t,
                     "config": {"workload": "direct 2x2 tables, row sums U{20..500} (BASELINE configs[3], one GPU's shard of "
                                            "100 M), log-space arithmetic; L2 flushed between steps", "tables": n},
                     "roofline": {"kernel": "fet_score", "bound": "hbm", "achieved": n * 24 / (ms_t * 1e-3) / 1e9, "peak": hbm,
                                  "unit": "GB/s", "frac": n * 24 / (ms_t * 1e-3) / 1e9 / hbm,
                                  "note": "fp64-pipe bound at this coverage (tail walks of 20-250 divide steps per table)"}}
    return out, (ch, hp, s_host)


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_css(sample_windows, steps=1, warmup=0):
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
                              CSS["wsize"], CSS["wstep"], av.size, bv.size, CSS["mct"], CSS["mcr"], 0, CSS["mds"],
                              checkers.dptr(s), checkers.dptr(p))
            dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return dict(seconds=times, windows=n, scores=s[:n], sub=(av, bv, sub["pos"], regend))


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


def css_config():
    return {"workload": "CSS scan, 450 Mb synthetic genome (BASELINE configs[2]): %d chromosomes x %d bp, %d SNPs each, "
                        "%d+%d individuals, wsize %d / wstep %d, classical MDS (mds=0), mcT=mcR=%d permutations; "
                        "genome inputs (1.44 GB float64) exceed L2, no flush" % (
                            CSS["chromosomes"], CSS["length"], CSS["nsnp"], CSS["asize"], CSS["bsize"], CSS["wsize"],
                            CSS["wstep"], CSS["mcr"]),
            "windows_per_step_per_gpu": CSS["chromosomes"] * (CSS["length"] // CSS["wstep"]), "sharding": "one genome per GPU (weak)"}


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
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1000.0, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": css_config(),
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


LARGE = dict(seed0=900, asize=500, bsize=500, windows=592, wsize=50_000, wstep=50_000, snps_per_window=167, mcr=1000)


def bench_large_cohort(lib_mod, args):
    """BASELINE configs[4] cohort (500+500 individuals, 50 kb windows, ~167 SNPs per window) on a slice of windows: the
    large-cohort kernels (Lanczos classical MDS; observed scores by a warp per window; permutation test with the between-group
    sums of 128 permutations as one tcgen05 u8 contraction, accumulators in tensor memory). Timed through
    the host scan entry with compact int8 inputs; the per-kernel split comes from the library's own CUDA events."""
    import ctypes as C
    import fpt_b200.api as api
    import fpt_b200.synth as synth
    lib = lib_mod.load()
    nwin = LARGE["windows"] if not args.small else 8
    regend = LARGE["wstep"] * nwin
    ch = synth.chromosome_fast(LARGE["seed0"], regend, LARGE["snps_per_window"] * nwin, LARGE["asize"], LARGE["bsize"], wstep=LARGE["wstep"])
    buf = C.create_string_buffer(8192)
    secs, prof = [], None
    lib.fpt_profile_enable(1)
    for it in range(3):                                          # first call pays the scratch allocation
        lib.fpt_profile_summary(buf, 8192)
        t0 = time.perf_counter()
        s, p, wr = api.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], LARGE["asize"], LARGE["bsize"], regend, LARGE["wsize"], LARGE["wstep"],
                                LARGE["mcr"], LARGE["mcr"], mds=0, seed=SEED)
        secs.append(time.perf_counter() - t0)
        lib.fpt_profile_summary(buf, 8192)
        prof = json.loads(buf.value.decode() or "{}")
    lib.fpt_profile_enable(0)
    scored = int((wr == 1).sum())
    dev_ms = sum(v["ms"] for v in prof.values())
    # the permutation kernel's contraction (csrc/fpt_css_perm_umma.cuh): per window ceil(mcr / 128) batches of a
    # 128 x Np x Kp u8 product per base-256 digit (4 digits), Np / Kp = m padded to 256 / 128
    m = LARGE["asize"] + LARGE["bsize"]
    npad, kpad, batches = -(-m // 256) * 256, -(-m // 128) * 128, -(-LARGE["mcr"] // 128)
    macs = float(scored) * batches * 128 * npad * kpad * 4
    perm_ms = prof.get("css_perm", {}).get("ms", 0.0)
    # kind::i8 rate on this part, from the ncu capture of this kernel (profiles/r1_css_perm_umma_raw.csv): the imma sub-pipe is
    # busy 512 cycles per 128 x 256 x 32 instruction = 2048 MAC/clk/SM (half the bf16 rate), times 148 SMs at the maximum SM clock
    i8_peak = 2048.0 * 2.0 * 148 * 1.965e9 / 1e12
    tensor = {"kernel": "css_perm (fpt_css_perm_umma_kernel)", "bound": "tensor", "unit": "TOP/s (u8 x u8 -> s32)",
              "achieved": 2.0 * macs / (perm_ms * 1e-3) / 1e12 if perm_ms > 0 else None,
              "peak": i8_peak,
              "peak_source": "tcgen05 kind::i8 issue rate measured with ncu on this kernel (512 pipe cycles per 128x256x32 MMA = 2048 MAC/clk/SM) x 148 SMs x 1965 MHz; "
                             "MEASURED_PEAKS.json has no 8-bit figure",
              "note": "whole-kernel time: the contraction is ~20 % of it and runs with the imma sub-pipe busy throughout (tensor-pipe bound); the rest is SIMT "
                      "work (surrogate distance pass, label shuffles, adjacent-pair sweep, membership rows)"}
    if tensor["achieved"] is not None:
        tensor["frac"] = tensor["achieved"] / tensor["peak"]
    return {"metric": "css_windows_per_sec_1000perms", "unit": "windows/s", "value": scored / (dev_ms * 1e-3), "roofline_tensor": tensor,
            "e2e": {"value": scored / min(secs[1:]), "unit": "windows/s", "api": "fpt_css_scan (host int8 codes)",
                    "h2d_bytes_per_step": int(ch["acodes"].nbytes + ch["bcodes"].nbytes + ch["pos"].nbytes)},
            "config": {"workload": "CSS scan, %d windows of BASELINE configs[4]: 500+500 individuals, 50 kb windows, ~%d SNPs per window, "
                                   "classical MDS, mcT=mcR=1000" % (nwin, LARGE["snps_per_window"]), "windows_scored": scored},
            "kernel_ms": {k: v["ms"] for k, v in prof.items()},
            "note": "device figure = windows / sum of the kernels' CUDA-event times of one scan; e2e = wall clock of the host call"}


def main():
    out = OneLineStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-fet", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-large", action="store_true", help="skip the 500+500 cohort slice")
    ap.add_argument("--small", action="store_true", help="tiny shapes: checks that the script runs, not a benchmark")
    ap.add_argument("--chromosomes", type=int, default=None, help="profiling aid: fewer chromosomes than the 21 of the workload")
    args = ap.parse_args()
    if args.small:
        CSS.update(chromosomes=2, length=2_000_000, nsnp=20_000)
        FET.update(length=10_000_000, nsnp=100_000)
        FET_TABLES.update(n=200_000)
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
    lib_mod.load().fpt_set_seed(SEED)
    css = bench_css(lib_mod, args, rank, world, dist)
    fet = None
    if rank == 0 and not args.skip_fet:
        fet, fet_sample = bench_fet(lib_mod, args)
    large = None
    if rank == 0 and not args.skip_large:
        large = bench_large_cohort(lib_mod, args)
    if world > 1:
        dist.barrier()
    if rank == 0:
        hbm, hbm_src = peaks()
        m = CSS["asize"] + CSS["bsize"]
        ms_step = css["ms"] / args.steps
        value = world * css["windows_per_step"] / (ms_step * 1e-3)
        e2e_val = world * css["windows_per_step"] / (css["ms_e2e"] / args.steps * 1e-3)
        prof = css["prof"]
        tot = sum(v["ms"] for v in prof.values())
        dom = max(prof, key=lambda k: prof[k]["ms"])
        per_launch_ms = prof[dom]["ms"] / prof[dom]["launches"]
        nout = CSS["length"] // CSS["wstep"]
        bytes_launch = css_kernel_bytes(dom, nout, CSS["nsnp"], m)
        ach = bytes_launch / (per_launch_ms * 1e-3) / 1e9
        line = {
            "metric": "css_windows_per_sec_1000perms", "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": css_config(),
            "windows_scored_per_step_per_gpu": css["scored"],
            "perm_rechecks": {"exact_rescorings": css["rechecks"],
                              "permutations_scored": css["scored"] * 1024 * (args.steps + args.warmup),
                              "note": "permutations whose integer surrogate score was too close to the observed score to decide "
                                      "`>=` and were re-scored in the reference's summation order"},
            "e2e": {"value": e2e_val, "unit": "windows/s", "h2d_bytes_per_step": css["h2d"], "d2h_bytes_per_step": css["d2h"],
                    "api": "fpt_css_compute (drop-in, host float64 arrays in pinned memory), one call per chromosome",
                    "windows_scored": css["scored_e2e"]},
            "gpu_launches": css["launches"],
            "roofline": {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                         "peak_source": hbm_src, "traffic": ncu_traffic(dom),
                         "share_of_step": {k: round(v["ms"] / tot, 4) for k, v in prof.items()},
                         "kernel_ms_per_launch": {k: v["ms"] / v["launches"] for k, v in prof.items()},
                         "note": "the permutation kernel is instruction-issue bound (48-bit LCG label shuffles, u8 IMMA between-group "
                                 "sums over a quantised distance matrix held in shared memory, exact fp64 re-scoring of near ties); "
                                 "it reads ~0.7 KB per window from HBM, so the HBM fraction is small by construction; see "
                                 "roofline_issue for the bound that applies"},
            "clocks": css["clocks"],
            "mds_variants": {"note": "same workload, chromosome 0 only, device-resident: mds=1 is SMACOF from 4 random starts, mds=2 is "
                                     "classical MDS followed by SMACOF (css.c:208-218)", **css["variants"]},
        }
        # issue-slot roofline of the dominant kernel: warp instructions of one launch (ncu capture of the same workload,
        # profiles/ncu_inst.json) over the live launch time, against 4 schedulers x 148 SMs x the sampled SM clock
        n_inst = ncu_traffic(dom, "inst")
        if n_inst and css["clocks"].get("sm_mhz"):
            issue_peak = 4 * 148 * css["clocks"]["sm_mhz"] * 1e6 / 1e9
            issue_ach = n_inst / (per_launch_ms * 1e-3) / 1e9
            line["roofline_issue"] = {"kernel": dom, "bound": "issue", "achieved": issue_ach, "peak": issue_peak,
                                      "unit": "Gwarp-inst/s", "frac": issue_ach / issue_peak,
                                      "peak_source": "1 warp-inst/clk/scheduler x 4 x 148 SMs x sampled SM clock",
                                      "inst_per_launch": n_inst}
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
            else:
                line["cpu_baseline"] = {"value": None, "unit": "windows/s", "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref not available"}
        if fet is not None:
            if world == 1 and not args.skip_cpu:
                cf = cpu_fet(200_000)
                if cf is not None:
                    ncores = os.cpu_count() or 1
                    ch, hp, s_host = fet_sample
                    n = cf["windows"]
                    fet["cpu_baseline"] = {"value": cf["snps"] / cf["seconds"], "unit": "SNPs/s", "cores": min(64, ncores), "threads": 64,
                                           "nproc": ncores, "kind": "reference",
                                           "sample": "first %d SNPs (%d windows) of the FET workload, reference threadcompute, %.1f s" % (
                                               cf["snps"], n, cf["seconds"]),
                                           "parity_vs_gpu": {"score_max_abs_diff": float(np.max(np.abs(s_host[:n - 8] - cf["scores"][:n - 8])))}}
            line["fet"] = fet
        if large is not None:
            line["large_cohort"] = large
        out.emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
