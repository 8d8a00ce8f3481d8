"""Extended host entry points of libfpt_b200.so (layer 2 of include/fpt_b200.h) over numpy arrays:
compact inputs, window ranges for sharding, explicit random-stream control and parity probes.
Nothing here computes on the CPU."""
import ctypes as C

import numpy as np

from . import _lib
from ._dropin import _arr
from ._lib import (FPT_SCAN_SERIAL, FPT_SCAN_THREADED, FPT_WIN_DISCARDED, FPT_WIN_EMPTY, FPT_WIN_SCORED, CssProbes,
                   FptError, Genotypes, ScanRange)

__all__ = ["FptError", "FPT_SCAN_SERIAL", "FPT_SCAN_THREADED", "FPT_WIN_EMPTY", "FPT_WIN_DISCARDED", "FPT_WIN_SCORED",
           "device_count", "set_device", "set_seed", "get_seed", "set_perm_mode", "get_perm_mode", "css_perm_rechecks", "window_state", "window_count", "fet_scan", "css_scan",
           "fet_per_snp", "fet_tables"]


def device_count():
    return _lib.load().fpt_device_count()


def set_device(device):
    _lib.check(_lib.load().fpt_set_device(int(device)))


def set_seed(seed):
    _lib.load().fpt_set_seed(int(seed) & 0xFFFFFFFFFFFFFFFF)


def get_seed():
    return int(_lib.load().fpt_get_seed())


def set_perm_mode(chain):
    """CSS label shuffles: False = independent per permutation (default), True = the reference's chained label array"""
    _lib.load().fpt_set_perm_mode(1 if chain else 0)


def get_perm_mode():
    return bool(_lib.load().fpt_get_perm_mode())


def set_lanczos_form(max_form):
    """Large-cohort MDS: highest matrix form the Lanczos product may stream (3 8-bit codes + blank list, 2 8-bit codes through a
    table, 1 16-bit count codes, 0 fp64 matrix)."""
    _lib.load().fpt_set_lanczos_form(int(max_form))


def set_perm_small_kernel(v):
    """Cohorts of 8..64, independent shuffles: 1 / True = fpt_css_perm3_kernel (default), 0 / False = the round-1 kernel."""
    _lib.load().fpt_set_perm_small_kernel(int(v))


def set_lanczos_threads(threads):
    """Large-cohort MDS on count codes: threads per CTA (256 default, 384 or 512). Tuning aid."""
    _lib.load().fpt_set_lanczos_threads(int(threads))


def set_mds_small_kernel(v):
    """Cohorts of 3..48, classical MDS: 1 / True = tridiagonalisation in registers (default), 0 / False = the shared-memory kernel."""
    _lib.load().fpt_set_mds_small_kernel(int(v))


def set_k4_mode(mode):
    """Large cohorts, genotype-distance matrix: 2 = tcgen05 u8 GEMM (default), 1 = popcounts, 0 = legacy fp64 matrix."""
    _lib.load().fpt_set_k4_mode(int(mode))


def k4_counts(a, b, pos, asize, bsize, regend, wsize, wstep, window, mode=2):
    """Parity probe: the m x m opposite-homozygote counts of one window as the tcgen05 GEMM (mode 2) or the popcount kernel (mode 1)
    produces them."""
    keep = []
    g = _genotypes(a, b, pos, asize, bsize, keep)
    r, n = _range(regend, wsize, wstep, FPT_SCAN_SERIAL, None, None, None, None, None, keep)
    m = asize + bsize
    out = np.zeros((m, m), dtype=np.int32)
    _lib.check(_lib.load().fpt_debug_k4_counts(C.byref(g), C.byref(r), int(mode), int(window), out.ctypes.data))
    return out


def set_perm_large_kernel(tensor_memory):
    """Large cohorts (m > 250): 1 / True = tcgen05 permutation kernel (default), 0 / False = the general kernel, 2 = tcgen05
    kernel with a coarse (10-bit) surrogate that forces many exact re-scorings. Same results in every mode."""
    _lib.load().fpt_set_perm_large_kernel(int(tensor_memory))


def css_perm_rechecks():
    return int(_lib.load().fpt_css_perm_rechecks())


def window_state(seed, window, stream=0):
    return int(_lib.load().fpt_window_state(int(seed) & 0xFFFFFFFFFFFFFFFF, int(window), int(stream)))


def window_count(regend, wsize, wstep):
    """Number of output slots of a scan: regend // wstep, the length the reference's callers allocate
    (statistics/FisherExactScoreStat.py:51-53)."""
    return regend // wstep


def _genotypes(a, b, pos, asize, bsize, keep):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    nsnp = pos.size
    g = Genotypes()
    if a.dtype == np.float64 and b.dtype == np.float64:
        g.avals, g.bvals = a.ctypes.data, b.ctypes.data
    elif a.dtype == np.int8 and b.dtype == np.int8:
        g.acodes, g.bcodes = a.ctypes.data, b.ctypes.data
    else:
        raise ValueError("genotypes must both be float64 (reference layout) or both int8 (compact codes)")
    if a.size != nsnp * asize or b.size != nsnp * bsize:
        raise ValueError("genotype arrays must hold nsnp*asize and nsnp*bsize values")
    g.pos, g.nsnp, g.asize, g.bsize = pos.ctypes.data, nsnp, int(asize), int(bsize)
    keep.extend([a, b, pos])
    return g


def _range(regend, wsize, wstep, semantics, window_begin, window_end, seed, states_resample, states_init, keep):
    r = ScanRange()
    r.regend, r.wsize, r.wstep, r.semantics = int(regend), int(wsize), int(wstep), int(semantics)
    r.window_begin = 0 if window_begin is None else int(window_begin)
    r.window_end = window_count(regend, wsize, wstep) if window_end is None else int(window_end)
    r.seed = (get_seed() if seed is None else int(seed)) & 0xFFFFFFFFFFFFFFFF
    n = r.window_end - r.window_begin
    for name, st in (("states_resample", states_resample), ("states_init", states_init)):
        if st is not None:
            st = np.ascontiguousarray(st, dtype=np.uint64)
            if st.size != n:
                raise ValueError("%s needs one state per window of the range" % name)
            keep.append(st)
            setattr(r, name, st.ctypes.data)
    return r, n


def _out(name, a, n):
    """caller-supplied output array: float64, 1-d, C-contiguous, writable, at least one entry per window of the range"""
    if a is None:
        return np.zeros(n)
    _arr(name, a, np.float64, writable=True)
    if a.size < n:
        raise ValueError("'%s' holds %d entries but the window range has %d" % (name, a.size, n))
    return a


def fet_scan(a, b, pos, asize, bsize, regend, wsize, wstep, perc, semantics=FPT_SCAN_SERIAL, window_begin=None,
             window_end=None, seed=None, states=None, scores=None, stddev=None):
    """Windowed FET scan. ``a``/``b``: float64 values or int8 codes, SNP-major; ``pos``: one position per SNP.
    Returns (scores, stddev, written) for windows [window_begin, window_end)."""
    keep = []
    g = _genotypes(a, b, pos, asize, bsize, keep)
    r, n = _range(regend, wsize, wstep, semantics, window_begin, window_end, seed, states, None, keep)
    scores, stddev = _out("scores", scores, n), _out("stddev", stddev, n)
    written = np.zeros(n, dtype=np.uint8)
    _lib.check(_lib.load().fpt_fet_scan(C.byref(g), C.byref(r), float(perc), scores.ctypes.data, stddev.ctypes.data,
                                        written.ctypes.data))
    return scores, stddev, written


def css_scan(a, b, pos, asize, bsize, regend, wsize, wstep, treshold, runs, drosophila=0, mds=0,
             semantics=FPT_SCAN_SERIAL, window_begin=None, window_end=None, seed=None, states_perm=None,
             states_init=None, probes=False, scores=None, p=None):
    """Windowed CSS scan. Returns (scores, p, written) and, with ``probes=True``, a dict of per-window
    diagnostics (status, X, evals, hits, nperm, smacof_iters, smacof_sigma)."""
    keep = []
    g = _genotypes(a, b, pos, asize, bsize, keep)
    r, n = _range(regend, wsize, wstep, semantics, window_begin, window_end, seed, states_perm, states_init, keep)
    m = asize + bsize
    scores, p = _out("scores", scores, n), _out("p", p, n)
    written = np.zeros(n, dtype=np.uint8)
    pr, out = None, None
    if probes:
        nruns = 4 if mds == 1 else (1 if mds == 2 else 0)
        out = {"status": np.zeros(n, dtype=np.uint8), "X": np.zeros((n, m, 2)), "evals": np.zeros((n, 3)),
               "hits": np.zeros(n, dtype=np.int32), "nperm": np.zeros(n, dtype=np.int32),
               "smacof_iters": np.zeros((n, max(nruns, 1)), dtype=np.int32), "smacof_sigma": np.zeros((n, max(nruns, 1)))}
        pr = CssProbes()
        for k, v in out.items():
            setattr(pr, k, v.ctypes.data)
    _lib.check(_lib.load().fpt_css_scan(C.byref(g), C.byref(r), int(treshold), int(runs), int(drosophila), int(mds),
                                        scores.ctypes.data, p.ctypes.data, written.ctypes.data,
                                        C.byref(pr) if pr is not None else None))
    return (scores, p, written, out) if probes else (scores, p, written)


def fet_per_snp(a, b, asize, bsize, want_tables=True, want_scores=True):
    """Per-SNP stage only: (tables[nsnp,4] int32, neglog10p[nsnp])."""
    keep = []
    a = np.ascontiguousarray(a)
    nsnp = a.size // asize
    g = _genotypes(a, b, np.zeros(nsnp, dtype=np.int32), asize, bsize, keep)
    tables = np.zeros((nsnp, 4), dtype=np.int32) if want_tables else None
    sc = np.zeros(nsnp) if want_scores else None
    _lib.check(_lib.load().fpt_fet_per_snp(C.byref(g), tables.ctypes.data if want_tables else None,
                                           sc.ctypes.data if want_scores else None))
    return tables, sc


def fet_tables(tables, force_log=False):
    """-log10 P for direct 2x2 tables, int32 [n,4] = (A major, A minor, B major, B minor)."""
    t = np.ascontiguousarray(tables, dtype=np.int32)
    if t.ndim != 2 or t.shape[1] != 4:
        raise ValueError("tables must be [n,4]")
    out = np.zeros(t.shape[0])
    _lib.check(_lib.load().fpt_fet_tables(t.ctypes.data, t.shape[0], int(bool(force_log)), out.ctypes.data))
    return out
