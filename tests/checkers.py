"""ctypes bindings for the two CPU checkers used by the tests (never by the product):

* ``oracle``  — oracle/liboracle.so, our plain-C restatement (oracle/fpt_oracle.c)
* ``ref_fet`` / ``ref_css`` — oracle/_ref/libref_{fisher,css}.so, the unmodified reference C
  compiled by oracle/Makefile (present in the build container and, as prebuilt files, on the GPU box)

The reference libraries export several non-prototyped ``inline`` functions (percentile, stress, ...);
their signatures are declared here by hand.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_i64_p = C.POINTER(C.c_int64)
c_u64_p = C.POINTER(C.c_uint64)
c_ushort_p = C.POINTER(C.c_ushort)


def _ensure_built():
    lib = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "fpt_oracle.c")
    if (not os.path.exists(lib)) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/statistics") and not os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libref_css.so")):
        subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=True, capture_output=True)


def dptr(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_int_p)


def load_oracle():
    _ensure_built()
    o = C.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))
    o.fpt_oracle_window_state.restype = C.c_uint64
    o.fpt_oracle_window_state.argtypes = [C.c_uint64, C.c_int64, C.c_int]
    o.fpt_oracle_nrand48.restype = C.c_long
    o.fpt_oracle_nrand48.argtypes = [c_u64_p]
    o.fpt_oracle_drand48.restype = C.c_double
    o.fpt_oracle_drand48.argtypes = [c_u64_p]
    o.fpt_oracle_randint.restype = C.c_long
    o.fpt_oracle_randint.argtypes = [C.c_long, c_u64_p]
    o.fpt_oracle_window_count.restype = C.c_int64
    o.fpt_oracle_window_count.argtypes = [C.c_int] * 3
    o.fpt_oracle_window_scheduled.restype = C.c_int
    o.fpt_oracle_window_scheduled.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]
    o.fpt_oracle_window_bounds.restype = None
    o.fpt_oracle_window_bounds.argtypes = [c_int_p, C.c_int64, C.c_int64, C.c_int, C.c_int, c_i64_p, c_i64_p]
    o.fpt_oracle_fetcount.restype = None
    o.fpt_oracle_fetcount.argtypes = [c_double_p, c_double_p, C.c_int64, C.c_int, C.c_int, c_int_p]
    o.fpt_oracle_binomial.restype = C.c_uint64
    o.fpt_oracle_binomial.argtypes = [C.c_uint64, C.c_uint64]
    o.fpt_oracle_fet_exact_domain.restype = C.c_int
    o.fpt_oracle_fet_exact_domain.argtypes = [c_int_p]
    for name in ("fpt_oracle_fet_exact", "fpt_oracle_fet_neglog10", "fpt_oracle_fet_neglog10_logmode"):
        getattr(o, name).restype = C.c_double
        getattr(o, name).argtypes = [c_int_p]
    o.fpt_oracle_percentile.restype = C.c_double
    o.fpt_oracle_percentile.argtypes = [c_double_p, C.c_int, C.c_double]
    o.fpt_oracle_fet_window.restype = None
    o.fpt_oracle_fet_window.argtypes = [c_double_p, C.c_int, C.c_double, C.c_int, C.c_uint64, c_double_p]
    o.fpt_oracle_fet_scan.restype = C.c_int
    o.fpt_oracle_fet_scan.argtypes = [c_double_p, c_double_p, c_int_p, c_int_p] + [C.c_int] * 6 + [
        C.c_double, c_double_p, c_double_p, C.c_int, C.c_uint64]
    o.fpt_oracle_fet_per_snp.restype = C.c_int
    o.fpt_oracle_fet_per_snp.argtypes = [c_double_p, c_double_p, C.c_int64, C.c_int, C.c_int, c_int_p, c_double_p]
    o.fpt_oracle_fet_tables.restype = None
    o.fpt_oracle_fet_tables.argtypes = [c_int_p, C.c_int64, c_double_p]
    o.fpt_oracle_compare_all.restype = None
    o.fpt_oracle_compare_all.argtypes = [c_double_p, c_double_p, C.c_int, C.c_int, C.c_int, c_double_p]
    o.fpt_oracle_compare_freq.restype = None
    o.fpt_oracle_compare_freq.argtypes = [c_double_p, c_double_p, C.c_int, c_double_p]
    o.fpt_oracle_fill_averages.restype = C.c_int
    o.fpt_oracle_fill_averages.argtypes = [c_double_p, C.c_int]
    o.fpt_oracle_cmds.restype = None
    o.fpt_oracle_cmds.argtypes = [c_double_p, C.c_int, c_double_p, c_double_p]
    o.fpt_oracle_smacof.restype = C.c_double
    o.fpt_oracle_smacof.argtypes = [c_double_p, C.c_int, c_double_p, C.c_int, C.c_double, c_int_p]
    o.fpt_oracle_smacof_margin.restype = C.c_double
    o.fpt_oracle_smacof_margin.argtypes = [c_double_p, C.c_int, c_double_p, C.c_int, C.c_double, c_int_p, c_double_p]
    o.fpt_oracle_smacof_runs.restype = C.c_double
    o.fpt_oracle_smacof_runs.argtypes = [c_double_p, C.c_int, c_double_p, C.c_int, C.c_int, C.c_double, c_u64_p]
    o.fpt_oracle_calc_dist.restype = None
    o.fpt_oracle_calc_dist.argtypes = [c_double_p, C.c_int, c_double_p]
    o.fpt_oracle_css.restype = C.c_double
    o.fpt_oracle_css.argtypes = [c_double_p, C.c_int, c_int_p, c_int_p, C.c_int, C.c_int]
    o.fpt_oracle_significance.restype = C.c_double
    o.fpt_oracle_significance.argtypes = [c_double_p, C.c_int, c_int_p, C.c_int, C.c_int, C.c_double, C.c_int,
                                          C.c_int, c_u64_p, c_int_p, c_int_p]
    o.fpt_oracle_lcg_skip.restype = C.c_uint64
    o.fpt_oracle_lcg_skip.argtypes = [C.c_uint64, C.c_uint64]
    o.fpt_oracle_significance_indep.restype = C.c_double
    o.fpt_oracle_significance_indep.argtypes = [c_double_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_uint64,
                                                c_int_p, c_int_p]
    o.fpt_oracle_set_perm_mode.restype = None
    o.fpt_oracle_set_perm_mode.argtypes = [C.c_int]
    o.fpt_oracle_css_window.restype = C.c_double
    o.fpt_oracle_css_window.argtypes = [c_double_p, c_double_p] + [C.c_int] * 7 + [C.c_uint64, C.c_uint64,
                                                                                  c_double_p, c_double_p, c_double_p]
    o.fpt_oracle_css_scan.restype = C.c_int
    o.fpt_oracle_css_scan.argtypes = [c_double_p, c_double_p, c_int_p, c_int_p] + [C.c_int] * 10 + [
        c_double_p, c_double_p, C.c_int, C.c_uint64]
    return o


def ref_available():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libref_fisher.so")) and \
        os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libref_css.so"))


def load_ref_fet():
    _ensure_built()
    r = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_fisher.so"))
    r.binomial.restype = C.c_ulong
    r.binomial.argtypes = [C.c_ulong, C.c_ulong]
    r.fet.restype = C.c_double
    r.fet.argtypes = [c_int_p, c_int_p]
    r.fet_p.restype = C.c_double
    r.fet_p.argtypes = [C.c_int] * 4
    r.fetcount.restype = None
    r.fetcount.argtypes = [c_int_p, c_double_p, c_double_p, C.c_int, C.c_int, C.c_int]
    r.shift_table.restype = None
    r.shift_table.argtypes = [c_int_p, c_int_p, C.c_int]
    r.create_table.restype = None
    r.create_table.argtypes = [c_int_p, C.c_int]
    r.min_idx.restype = C.c_int
    r.min_idx.argtypes = [c_int_p, C.c_int]
    r.percentile.restype = C.c_double
    r.percentile.argtypes = [c_double_p, C.c_int, C.c_double]
    r.std.restype = C.c_double
    r.std.argtypes = [c_double_p, C.c_int]
    r.mean.restype = C.c_double
    r.mean.argtypes = [c_double_p, C.c_int]
    r.calc_std.restype = C.c_double
    r.calc_std.argtypes = [c_double_p, c_double_p, c_double_p, C.c_int, C.c_int, C.c_double, c_ushort_p]
    r.random_int_nrand48.restype = C.c_long
    r.random_int_nrand48.argtypes = [C.c_long, c_ushort_p]
    r.fisher_exact_test.restype = None
    r.fisher_exact_test.argtypes = [c_double_p, c_double_p, c_double_p, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p,
                                    c_double_p, c_double_p, C.c_int, c_double_p, c_ushort_p, C.c_double]
    r.get_population_size.restype = C.c_int
    r.get_population_size.argtypes = [c_int_p]
    r.slide_right.restype = None
    r.slide_right.argtypes = [c_int_p, c_int_p, C.c_int, C.c_int, C.c_int]
    sig = [c_double_p, c_double_p, c_int_p, c_int_p] + [C.c_int] * 6 + [C.c_double, c_double_p, c_double_p]
    r.compute.restype = None
    r.compute.argtypes = sig
    r.threadcompute.restype = None
    r.threadcompute.argtypes = sig
    return r


class RefMatrix:
    """double** view over one contiguous block, the layout of css.c:375-404 allocate_matrix."""

    def __init__(self, rows, cols, init=None):
        self.a = np.zeros((rows, cols), dtype=np.float64) if init is None else np.ascontiguousarray(init, dtype=np.float64)
        rows, cols = self.a.shape
        self.rowptr = (c_double_p * rows)()
        base = self.a.ctypes.data
        for i in range(rows):
            self.rowptr[i] = C.cast(base + i * cols * 8, c_double_p)

    @property
    def pp(self):
        return C.cast(self.rowptr, C.POINTER(c_double_p))


def load_ref_css():
    _ensure_built()
    r = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_css.so"))
    pp = C.POINTER(c_double_p)
    r.compare_all.restype = None
    r.compare_all.argtypes = [c_double_p, c_double_p, C.c_int, C.c_int, C.c_int, pp]
    r.compare_freq.restype = None
    r.compare_freq.argtypes = [c_double_p, c_double_p, C.c_int, pp]
    r.fill_averages.restype = C.c_int
    r.fill_averages.argtypes = [pp, C.c_int]
    r.matrix_mult.restype = None
    r.matrix_mult.argtypes = [pp, pp, pp, C.c_int, C.c_int, C.c_int]
    r.setup_z_matrix.restype = None
    r.setup_z_matrix.argtypes = [pp, C.c_int]
    r.cmds.restype = None
    r.cmds.argtypes = [pp, pp, C.c_int, C.c_int, pp, pp, pp, pp, pp]
    r.calc_dist.restype = None
    r.calc_dist.argtypes = [pp, pp, C.c_int]
    r.css.restype = C.c_double
    r.css.argtypes = [pp, c_int_p, c_int_p, C.c_int, C.c_int]
    r.stress.restype = C.c_double
    r.stress.argtypes = [pp, pp, C.c_int]
    r.random_shuffle.restype = None
    r.random_shuffle.argtypes = [c_int_p, C.c_int, c_ushort_p]
    r.significance_treshold.restype = C.c_double
    r.significance_treshold.argtypes = [pp, c_int_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_ushort_p]
    r.smacof.restype = C.c_double
    r.smacof.argtypes = [pp, C.c_int, C.c_int, pp, pp, pp, pp, C.c_int, C.c_double]
    r.smacof_runs.restype = None
    r.smacof_runs.argtypes = [pp, C.c_int, C.c_int, pp, pp, pp, pp, pp, C.c_int, C.c_int, C.c_double]
    r.cluster_separation_scorer.restype = C.c_double
    r.cluster_separation_scorer.argtypes = [pp, c_double_p, c_double_p, c_int_p, c_int_p, C.c_int, C.c_int, C.c_int,
                                            pp, C.c_int, C.c_int, pp, pp, pp, pp, pp, pp, pp]
    sig = [c_double_p, c_double_p, c_int_p, c_int_p] + [C.c_int] * 10 + [c_double_p, c_double_p]
    r.compute.restype = None
    r.compute.argtypes = sig
    r.threadcompute.restype = None
    r.threadcompute.argtypes = sig
    return r


def state_to_ushort3(state48):
    """packed 48-bit LCG state -> the unsigned short[3] the reference passes to nrand48"""
    arr = (C.c_ushort * 3)(state48 & 0xFFFF, (state48 >> 16) & 0xFFFF, (state48 >> 32) & 0xFFFF)
    return arr


def ushort3_to_state(arr):
    return int(arr[0]) | (int(arr[1]) << 16) | (int(arr[2]) << 32)


def seed48(state48):
    """set glibc's global drand48 state (used by the reference's smacof_runs, css.c:863-864)"""
    libc = C.CDLL(None)
    libc.seed48.restype = c_ushort_p
    libc.seed48.argtypes = [c_ushort_p]
    libc.seed48(state_to_ushort3(state48))


def silence_stdout():
    """context manager: the reference printf()s progress lines on every compute call"""
    import contextlib
    import sys

    @contextlib.contextmanager
    def _cm():
        sys.stdout.flush()
        libc = C.CDLL(None)
        libc.fflush(None)
        saved = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            yield
        finally:
            libc.fflush(None)
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
    return _cm()


def fet_exact_rule(a, b, c, d):
    """The reference's two-tailed rule (fisher/cFisher.c:245-340: first tail towards the minimum cell, strict `<` on the second
    tail, doubling on equal margins) in exact rational arithmetic; returns P as a Fraction."""
    from fractions import Fraction
    from math import comb
    R1, R2, C1, C2 = a + b, c + d, a + c, b + d
    n = R1 + R2
    cw = [a, b, d, c]
    at = cw.index(min(cw))
    x_dir = -1 if at in (0, 2) else 1                     # minimum cell is a or d: x = cell(0,0) decreases
    lo, hi = max(0, C1 - R2), min(R1, C1)
    pm = lambda x: Fraction(comb(R1, x) * comb(R2, C1 - x), comb(n, C1))
    P0 = pm(a)
    xs = range(lo, a + 1) if x_dir < 0 else range(a, hi + 1)
    P = sum(pm(x) for x in xs)
    if R1 == R2 or C1 == C2:
        P = 2 * P
    else:
        other = range(hi, a, -1) if x_dir < 0 else range(lo, a)
        for x in other:
            if pm(x) < P0:
                P += pm(x)
            else:
                break
    return min(P, Fraction(1))
