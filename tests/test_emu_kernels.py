"""The product's CUDA kernels (csrc/fpt_*.cuh) executed on the CPU through tests/emu/cuda_emu.h — one pthread per
CUDA thread, real barriers, emulated warp collectives — and compared with the oracle. This keeps kernel LOGIC
(indexing, barriers, scans, skip-ahead streams, rejection repair) under test where no GPU exists; the same source
is what nvcc compiles for sm_100a. Sizes are tiny because every __syncthreads is a pthread barrier."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from checkers import dptr, iptr

HERE = os.path.dirname(os.path.abspath(__file__))
CODES = np.array([3.0, -3.0, 0.0, -10000.0])


@pytest.fixture(scope="module")
def emu():
    # FPT_EMU_LIB: a prebuilt variant of the emulation library, e.g. the ThreadSanitizer / AddressSanitizer builds that
    # profiles/emu_sanitizers.sh runs this suite under (race and bounds checks of the kernel logic without a GPU tool)
    override = os.environ.get("FPT_EMU_LIB")
    if override:
        return C.CDLL(override)
    subprocess.run([os.path.join(HERE, "emu", "build.sh")], check=True, capture_output=True)
    return C.CDLL(os.path.join(HERE, "emu", "libfpt_emu.so"))


@pytest.fixture(scope="module")
def emu_exact_stress():
    """the same kernels with the SMACOF order bound scaled by 1e12: every stopping decision takes the path that re-sums the
    stress in the reference's order (csrc/fpt_css.cuh, fpt_css_smacof)"""
    override = os.environ.get("FPT_EMU_LIB_EXACT")           # sanitizer builds, see the emu fixture
    if override:
        return C.CDLL(override)
    out = os.path.join(HERE, "emu", "libfpt_emu_exact.so")
    subprocess.run([os.path.join(HERE, "emu", "build.sh"), out, "-DFPT_SMACOF_BOUND_SCALE=1e12"], check=True, capture_output=True)
    return C.CDLL(out)


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def ll(v):
    return C.c_longlong(int(v))


def test_skip_ahead_matches_stepping(emu, oracle):
    emu.emu_lcg_skip.restype = C.c_uint64
    emu.emu_lcg_skip.argtypes = [C.c_uint64, C.c_uint64]
    emu.emu_window_state.restype = C.c_uint64
    emu.emu_window_state.argtypes = [C.c_uint64, C.c_longlong, C.c_int]
    for seed, w, stream in ((1, 0, 0), (20261018, 123456, 1), (2 ** 64 - 1, 2 ** 31, 0)):
        assert emu.emu_window_state(seed, w, stream) == oracle.fpt_oracle_window_state(seed, w, stream)
    s = C.c_uint64(0x1234ABCD5678)
    for n in (0, 1, 2, 39, 1000, 12345):
        t = C.c_uint64(0x1234ABCD5678)
        for _ in range(n):
            oracle.fpt_oracle_nrand48(C.byref(t))
        assert emu.emu_lcg_skip(s.value, n) == t.value


def test_fet_count_and_score_kernels(emu, oracle):
    rng = np.random.default_rng(1)
    asize, bsize, S = 20, 17, 333
    av, bv = rng.choice(CODES, S * asize, p=[.45, .3, .23, .02]), rng.choice(CODES, S * bsize, p=[.3, .45, .23, .02])
    tab = np.zeros((S, 4), dtype=np.int32)
    emu.emu_fet_count_f64(dptr(av), dptr(bv), ll(S), asize, bsize, 64, 3, iptr(tab))
    tab_o, sc_o = np.zeros((S, 4), dtype=np.int32), np.zeros(S)
    oracle.fpt_oracle_fet_per_snp(dptr(av), dptr(bv), S, asize, bsize, iptr(tab_o), dptr(sc_o))
    assert np.array_equal(tab, tab_o)
    a8 = np.where(av == -10000, -128, av).astype(np.int8)
    b8 = np.where(bv == -10000, -128, bv).astype(np.int8)
    tab8 = np.zeros((S, 4), dtype=np.int32)
    emu.emu_fet_count_i8(vp(a8), vp(b8), ll(S), asize, bsize, 64, 3, iptr(tab8))
    assert np.array_equal(tab8, tab_o)
    sc = np.zeros(S)
    emu.emu_fet_score(iptr(tab), ll(S), asize + bsize, 1, 0, 2, dptr(sc))
    assert np.array_equal(sc, sc_o)                  # same libm on the host: bit-identical
    # log mode, coverage up to 500, log-factorials read from "global" memory
    n1, n2 = rng.integers(20, 501, 500), rng.integers(20, 501, 500)
    f = np.clip(rng.beta(.5, .5, 500), .02, .98)
    a, c = rng.binomial(n1, f), rng.binomial(n2, np.clip(f + rng.normal(0, .1, 500), .01, .99))
    T = np.stack([a, n1 - a, c, n2 - c], 1).astype(np.int32).copy()
    so, se = np.zeros(500), np.zeros(500)
    oracle.fpt_oracle_fet_tables(iptr(T), 500, dptr(so))
    assert emu.emu_fet_maxn(iptr(T), ll(500)) == T.sum(1).max()
    emu.emu_fet_score(iptr(T), ll(500), int(T.sum(1).max()), 0, 0, 2, dptr(se))
    # the kernel's log-mode walk multiplies by a Newton reciprocal where the oracle divides: equal to rounding
    np.testing.assert_allclose(se, so, rtol=1e-10, atol=1e-12)
    # the tile-sorting form (ragged last tile, exact-mode and log-mode tables mixed): the same scores, bit for bit, at the same indices
    nbig = 2 * 1024 + 77
    n1, n2 = rng.integers(2, 501, nbig), rng.integers(2, 501, nbig)
    f = np.clip(rng.beta(.5, .5, nbig), .02, .98)
    a, c = rng.binomial(n1, f), rng.binomial(n2, np.clip(f + rng.normal(0, .1, nbig), .01, .99))
    Tb = np.stack([a, n1 - a, c, n2 - c], 1).astype(np.int32).copy()
    s_plain, s_sorted = np.zeros(nbig), np.full(nbig, -7.0)
    emu.emu_fet_score(iptr(Tb), ll(nbig), int(Tb.sum(1).max()), 1, 0, 2, dptr(s_plain))
    emu.emu_fet_score_sorted(iptr(Tb), ll(nbig), int(Tb.sum(1).max()), 1, 0, 2, dptr(s_sorted))
    assert np.array_equal(s_plain, s_sorted)


@pytest.mark.parametrize("threaded", [0, 1])
def test_window_table_kernel(emu, oracle, threaded):
    rng = np.random.default_rng(2)
    for regend, wsize, wstep, S in ((60000, 2500, 500, 400), (51000, 1000, 1000, 300), (33333, 700, 300, 300), (20000, 2500, 500, 100)):
        pos = np.sort(rng.choice(regend, size=S, replace=False)).astype(np.int32)
        nwin = regend // wstep
        for wbase in (0, 7):
            n = nwin - wbase
            wl, wr = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
            emu.emu_window_table(iptr(pos), ll(S), ll(wbase), ll(n), regend, wsize, wstep, threaded, iptr(wl), iptr(wr))
            for j in range(n):
                w = wbase + j
                l, r = C.c_int64(), C.c_int64()
                oracle.fpt_oracle_window_bounds(iptr(pos), S, w, wsize, wstep, C.byref(l), C.byref(r))
                if oracle.fpt_oracle_window_scheduled(w, regend, wsize, wstep, threaded):
                    assert (wl[j], wr[j]) == (l.value, r.value)
                else:
                    assert wl[j] == wr[j]


@pytest.mark.parametrize("use_hist", [1, 0])
def test_fet_window_kernel(emu, oracle, use_hist):
    rng = np.random.default_rng(3)
    asize = bsize = 10
    regend, wsize, wstep, S = 30000, 2500, 500, 260
    pos = np.sort(rng.choice(regend, size=S, replace=False)).astype(np.int32)
    av, bv = rng.choice(CODES, S * asize, p=[.45, .3, .23, .02]), rng.choice(CODES, S * bsize, p=[.3, .45, .23, .02])
    apos, bpos = np.repeat(pos, asize).astype(np.int32), np.repeat(pos, bsize).astype(np.int32)
    n = regend // wstep
    s_o, d_o = np.zeros(n), np.zeros(n)
    oracle.fpt_oracle_fet_scan(dptr(av), dptr(bv), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, av.size, bv.size, 0.95,
                               dptr(s_o), dptr(d_o), 0, 777)
    snp = np.zeros(S)
    oracle.fpt_oracle_fet_per_snp(dptr(av), dptr(bv), S, asize, bsize, None, dptr(snp))
    wl, wr = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    mx = emu.emu_window_table(iptr(pos), ll(S), ll(0), ll(n), regend, wsize, wstep, 0, iptr(wl), iptr(wr))
    s_e, d_e, fl = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.uint8)
    emu.emu_fet_window(dptr(snp), iptr(wl), iptr(wr), ll(0), ll(n), C.c_double(0.95), C.c_uint64(777), None, int(mx), use_hist, 3,
                       dptr(s_e), dptr(d_e), vp(fl))
    assert np.array_equal(s_e, s_o) and np.array_equal(d_e, d_o)
    assert np.array_equal(fl == 1, (wr - wl) > 0)


def _css_input(seed, asize, bsize, S, L):
    rng = np.random.default_rng(seed)
    pos = np.sort(rng.choice(L, size=S, replace=False)).astype(np.int32)
    f = np.clip(rng.beta(.5, .5, S), .02, .98)
    fb = np.where((np.arange(S) // 40) % 3 == 0, 1 - f, f)
    enc = np.array([3., 0., -3.])
    av, bv = enc[rng.binomial(2, f[:, None], size=(S, asize))], enc[rng.binomial(2, fb[:, None], size=(S, bsize))]
    av[rng.random(av.shape) < .02] = -10000
    bv[rng.random(bv.shape) < .02] = -10000
    return pos, av.ravel().copy(), bv.ravel().copy()


def _pack(emu, av, bv, S, asize, bsize):
    m = asize + bsize
    planes = np.zeros(((S + 31) // 32) * 2 * m, dtype=np.uint32)
    emu.emu_css_pack_f64(dptr(av), dptr(bv), ll(S), asize, bsize, 2, 3, vp(planes))
    return planes


def test_css_pack_kernel(emu):
    asize, bsize, S = 5, 6, 100
    pos, av, bv = _css_input(5, asize, bsize, S, 5000)
    m = asize + bsize
    P = _pack(emu, av, bv, S, asize, bsize).reshape(-1, 2, m)
    G = np.concatenate([av.reshape(S, asize), bv.reshape(S, bsize)], 1)
    for k in range(S):
        w, b = divmod(k, 32)
        assert np.array_equal((P[w, 0] >> b) & 1, (G[k] == 3).astype(np.uint32))
        assert np.array_equal((P[w, 1] >> b) & 1, (G[k] == -3).astype(np.uint32))
    a8 = np.where(av == -10000, -128, av).astype(np.int8)
    b8 = np.where(bv == -10000, -128, bv).astype(np.int8)
    P8 = np.zeros_like(P.reshape(-1))
    emu.emu_css_pack_i8(vp(a8), vp(b8), ll(S), asize, bsize, 2, 3, vp(P8))
    assert np.array_equal(P8, P.reshape(-1))


@pytest.mark.parametrize("kernel,asize,bsize,S,L", [("smem", 6, 5, 220, 20000), ("reg", 6, 5, 220, 20000), ("reg", 1, 2, 120, 8000),
                                                   ("reg", 16, 16, 200, 10000), ("reg", 20, 20, 200, 10000), ("reg", 17, 20, 150, 8000),
                                                   ("reg", 25, 23, 220, 9000), ("reg", 20, 20, 2500, 9000)])
def test_css_mds_kernels(emu, oracle, kernel, asize, bsize, S, L):
    """phase A + phase B of the one-warp classical MDS against the oracle's cmds (css.c:505-560): the shared-memory
    tridiagonalisation and the register one (fpt_css_eig_reg.cuh) in each of its three paddings (32, 40, 48), a 3-individual
    cohort, and dense windows that span several 32-SNP words"""
    wsize, wstep = 2500, 500
    m = asize + bsize
    pos, av, bv = _css_input(6, asize, bsize, S, L)
    planes = _pack(emu, av, bv, S, asize, bsize)
    n = L // wstep
    wl, wr = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    emu.emu_window_table(iptr(pos), ll(S), ll(0), ll(n), L, wsize, wstep, 0, iptr(wl), iptr(wr))
    X, ev, st = np.zeros((n, m, 2)), np.zeros((n, 3)), np.zeros(n, dtype=np.uint8)
    if kernel == "reg":
        emu.emu_css_mds_warp_reg(vp(planes), m, iptr(wl), iptr(wr), ll(n), 2, dptr(X), dptr(ev), vp(st))
    else:
        emu.emu_css_mds_warp(vp(planes), None, m, iptr(wl), iptr(wr), ll(n), 4, 2, 3, dptr(X), dptr(ev), vp(st))
    scored = 0
    for w in range(n):
        l, r = int(wl[w]), int(wr[w])
        if r <= l:
            assert st[w] == 0
            continue
        D = np.zeros((m, m))
        oracle.fpt_oracle_compare_all(dptr(av[l * asize:r * asize].copy()), dptr(bv[l * bsize:r * bsize].copy()), asize, bsize, r - l, dptr(D))
        keep = oracle.fpt_oracle_fill_averages(dptr(D), m)
        assert st[w] == (2 if keep else 1)
        if not keep:
            continue
        Xo, evo = np.zeros((m, 2)), np.zeros(3)
        oracle.fpt_oracle_cmds(dptr(D), m, dptr(Xo), dptr(evo))
        assert np.allclose(ev[w, :2], evo[:2], rtol=1e-10, atol=1e-10 * abs(evo[0]))
        do, dg = np.zeros((m, m)), np.zeros((m, m))
        oracle.fpt_oracle_calc_dist(dptr(Xo), m, dptr(do))
        oracle.fpt_oracle_calc_dist(dptr(X[w].copy()), m, dptr(dg))
        if np.isfinite(do).all() and np.isfinite(dg).all() and evo[1] - evo[2] > 1e-8 * evo[0]:
            assert np.allclose(do, dg, rtol=1e-9, atol=1e-10 * do.max())
            scored += 1
    assert scored >= (10 if m > 3 else 3)


_LARGE_SHAPES = [(6, 5, 220, 20000), (30, 34, 170, 7000), (1, 1, 60, 6000), (9, 8, 3000, 12000)]


@pytest.mark.parametrize("shape,route", [(sh, r) for sh in _LARGE_SHAPES for r in ("legacy", "codes", "codes_table")
                                         if not (r == "codes_table" and sh[0] in (30, 1))])
def test_css_mds_large_cohort_kernel(emu, oracle, shape, route):
    """the Lanczos kernel used beyond the one-warp path, run here on small and medium cohorts (it is size-agnostic): early
    stop by the residual test (m = 64), complete Krylov space (m = 11, 2), windows discarded by fill_averages"""
    asize, bsize, S, L = shape
    wsize, wstep = 2500, 500
    m = asize + bsize
    pos, av, bv = _css_input(6, asize, bsize, S, L)
    planes = _pack(emu, av, bv, S, asize, bsize)
    n = L // wstep
    wl, wr = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    emu.emu_window_table(iptr(pos), ll(S), ll(0), ll(n), L, wsize, wstep, 0, iptr(wl), iptr(wr))
    X, ev, st, steps = np.zeros((n, m, 2)), np.zeros((n, 3)), np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.int32)
    if route == "legacy":      # fp64 dissimilarity matrix in global memory, converted to codes in place where it qualifies
        emu.emu_css_mds_large(vp(planes), None, m, iptr(wl), iptr(wr), ll(n), 2, 64, 2, dptr(X), dptr(ev), vp(st), iptr(steps))
    else:                      # count codes from the popcount form of the genotype GEMM (u8, and u16 beyond 255 SNPs per window)
        # "codes": squares computed arithmetically, blanks kept as a list (windows with too many blanks fall back to the table);
        # "codes_table": every square through the shared-memory table
        emu.emu_css_mds_codes(vp(planes), m, iptr(wl), iptr(wr), ll(n), 64, 2, dptr(X), dptr(ev), vp(st), iptr(steps), 1 if route == "codes" else 0)
    scored = 0
    for w in range(n):
        l, r = int(wl[w]), int(wr[w])
        if r <= l:
            assert st[w] == 0
            continue
        D = np.zeros((m, m))
        oracle.fpt_oracle_compare_all(dptr(av[l * asize:r * asize].copy()), dptr(bv[l * bsize:r * bsize].copy()), asize, bsize, r - l, dptr(D))
        keep = oracle.fpt_oracle_fill_averages(dptr(D), m)
        assert st[w] == (2 if keep else 1)
        if not keep:
            continue
        assert 1 <= steps[w] <= m
        Xo, evo = np.zeros((m, 2)), np.zeros(3)
        oracle.fpt_oracle_cmds(dptr(D), m, dptr(Xo), dptr(evo))
        assert np.allclose(ev[w, :2], evo[:2], rtol=1e-9, atol=1e-10 * abs(evo[0]))
        do, dg = np.zeros((m, m)), np.zeros((m, m))
        oracle.fpt_oracle_calc_dist(dptr(Xo), m, dptr(do))
        oracle.fpt_oracle_calc_dist(dptr(X[w].copy()), m, dptr(dg))
        if np.isfinite(do).all() and np.isfinite(dg).all() and (m < 3 or evo[1] - evo[2] > 1e-6 * evo[0]):
            assert np.allclose(do, dg, rtol=1e-6, atol=1e-8 * do.max())
            scored += 1
    assert scored >= (10 if m > 2 else 5)
    if m == 64:
        assert steps[st == 2].max() < m          # the residual test stops well before the Krylov space is complete


@pytest.mark.parametrize("exact_path", [0, 1])
@pytest.mark.parametrize("mds", [1, 2])
def test_css_smacof_and_perm_kernels(emu, emu_exact_stress, oracle, mds, exact_path):
    if exact_path:
        emu = emu_exact_stress
    asize, bsize, S, L, wsize, wstep = 4, 4, 90, 6000, 2500, 500
    m = asize + bsize
    pos, av, bv = _css_input(7, asize, bsize, S, L)
    apos, bpos = np.repeat(pos, asize).astype(np.int32), np.repeat(pos, bsize).astype(np.int32)
    planes = _pack(emu, av, bv, S, asize, bsize)
    n = L // wstep
    wl, wr = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    emu.emu_window_table(iptr(pos), ll(S), ll(0), ll(n), L, wsize, wstep, 0, iptr(wl), iptr(wr))
    X, ev, st = np.zeros((n, m, 2)), np.zeros((n, 3)), np.zeros(n, dtype=np.uint8)
    seed = 99
    if mds == 2:
        emu.emu_css_mds_warp(vp(planes), None, m, iptr(wl), iptr(wr), ll(n), 4, 2, 2, dptr(X), dptr(ev), vp(st))
        X0 = X.copy()
        Xr, sg, it = np.zeros((n, 1, m, 2)), np.zeros(n), np.zeros(n, dtype=np.int32)
        emu.emu_css_smacof(vp(planes), None, m, iptr(wl), iptr(wr), ll(0), ll(n), 4, 1, 3, 1, 0, C.c_uint64(seed), None, 300,
                           C.c_double(1e-6), dptr(X), dptr(Xr), dptr(sg), iptr(it), vp(st))
        X = Xr[:, 0].copy()
    else:
        Xr, sg, it = np.zeros((n, 4, m, 2)), np.zeros(n * 4), np.zeros(n * 4, dtype=np.int32)
        emu.emu_css_smacof(vp(planes), None, m, iptr(wl), iptr(wr), ll(0), ll(n), 4, 1, 3, 4, 1, C.c_uint64(seed), None, 300,
                           C.c_double(1e-6), None, dptr(Xr), dptr(sg), iptr(it), vp(st))
        emu.emu_css_pick(dptr(Xr), dptr(sg), m, 4, ll(n), vp(st), dptr(X))
    # iteration counts and stresses are the reference's (css.c:907-938): the stopping rule is decided on running sums in the
    # reference's order wherever the order of the additions could matter
    nruns = 4 if mds == 1 else 1
    checked = 0
    for w in range(n):
        if st[w] != 2:
            continue
        l, r = int(wl[w]), int(wr[w])
        D = np.zeros((m, m))
        oracle.fpt_oracle_compare_all(dptr(av[l * asize:r * asize].copy()), dptr(bv[l * bsize:r * bsize].copy()), asize, bsize, r - l, dptr(D))
        assert oracle.fpt_oracle_fill_averages(dptr(D), m)
        state = C.c_uint64(oracle.fpt_oracle_window_state(seed, w, 1))
        for run in range(nruns):
            if mds == 1:
                Xs = np.array([oracle.fpt_oracle_drand48(C.byref(state)) for _ in range(2 * m)]).reshape(m, 2)
            else:
                Xs = X0[w].copy()
            if not np.isfinite(Xs).all():
                continue
            k = C.c_int(0)
            sig = oracle.fpt_oracle_smacof(dptr(D), m, dptr(Xs), 300, 1e-6, C.byref(k))
            assert it[w * nruns + run] == k.value
            assert sg[w * nruns + run] == sig
            assert np.array_equal(Xr.reshape(n, nruns, m, 2)[w, run], Xs)
            checked += 1
    assert checked >= 5
    sc, p, hits, nn = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    emu.emu_css_perm(dptr(X), m, asize, bsize, ll(0), ll(n), vp(st), 7, 150, C.c_uint64(seed), None, 1, 1, 64, 2, 0, 1, dptr(sc), dptr(p),
                     iptr(hits), iptr(nn))
    so, po = np.zeros(n), np.zeros(n)
    oracle.fpt_oracle_set_perm_mode(1)
    oracle.fpt_oracle_css_scan(dptr(av), dptr(bv), iptr(apos), iptr(bpos), 0, L, wsize, wstep, av.size, bv.size, 7, 150, 0, mds,
                               dptr(so), dptr(po), 0, seed)
    oracle.fpt_oracle_set_perm_mode(0)
    scored = st == 2
    assert np.array_equal(scored, po != 0) and scored.sum() >= 5
    ok = scored & np.isfinite(so)
    np.testing.assert_allclose(sc[ok], so[ok], rtol=1e-6, atol=1e-9)
    if mds == 1:
        assert np.array_equal(sc[ok], so[ok])        # random starts: the whole chain is the reference's arithmetic order
        assert np.array_equal(p[ok], po[ok])


@pytest.mark.parametrize("chain", [0, 1])
def test_css_perm_kernels_chunks_early_stop_and_scratch_paths(emu, oracle, chain):
    """both permutation kernels, both shuffle modes: several chunks per window (runs > permutations per chunk), early
    stop inside a chunk, 16-bit labels, global-memory scratch, the integer surrogate with forced ties (rechecks)"""
    import math
    emu.emu_css_perm2.restype = C.c_ulonglong
    rng = np.random.default_rng(8)
    for asize, bsize, n in ((5, 6, 6), (2, 2, 4), (1, 4, 3)):
        m = asize + bsize
        X = rng.normal(size=(n, m, 2))
        X[:, :asize, 0] += np.linspace(0.0, 1.5, n)[:, None]      # windows from "no separation" to "clear separation"
        if m == 4:
            X[1] = np.round(X[1])                                  # exact ties between permuted and observed scores
        st = np.full(n, 2, dtype=np.uint8)
        states = (np.arange(n, dtype=np.uint64) * 7919 + 13)
        if m == 11:
            # window 0: the very first draw (n = 11) comes out as 2^31 - 1 > limit, i.e. is rejected and redrawn; window 1:
            # the same for the first draw of permutation 3 (independent mode) -> optimistic shuffle, replay, resync
            A, Cc, M48 = 0x5DEECE66D, 0xB, (1 << 48) - 1
            ainv = pow(A, -1, 1 << 48)
            target = (0x7FFFFFFF << 17) | 0x1ABCD
            states[0] = ((target - Cc) * ainv) & M48
            s = states[0].item()
            assert ((s * A + Cc) & M48) >> 17 == 0x7FFFFFFF
            back = (target - Cc) * ainv & M48
            for _ in range(3 * (m - 1)):
                back = ((back - Cc) * ainv) & M48
            states[1] = back
        qbits = min(15 if 8 <= m <= 64 else 22, int(math.floor(math.log2(2 ** 31 / (min(asize, bsize) * m + 1)))))
        for tres, runs in ((5, 300), (1000, 100), (1, 70)):
            want_p, want_h, want_n = [], [], []
            for w in range(n):
                dist = np.zeros((m, m))
                oracle.fpt_oracle_calc_dist(dptr(X[w].copy()), m, dptr(dist))
                tr = np.arange(m, dtype=np.int32)
                score = oracle.fpt_oracle_css(dptr(dist), m, iptr(tr), iptr(tr[asize:]), asize, bsize)
                h, nn = C.c_int(), C.c_int()
                if chain:
                    s64 = C.c_uint64(int(states[w]))
                    want_p.append(oracle.fpt_oracle_significance(dptr(dist), m, iptr(tr), asize, bsize, score, tres, runs, C.byref(s64), C.byref(h), C.byref(nn)))
                else:
                    want_p.append(oracle.fpt_oracle_significance_indep(dptr(dist), m, asize, bsize, score, tres, runs, int(states[w]), C.byref(h), C.byref(nn)))
                want_h.append(h.value)
                want_n.append(nn.value)
            for dist_smem, tracks_smem, wide in ((1, 1, 0), (0, 0, 1), (1, 0, 0)):
                sc, p, hits, nn = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
                emu.emu_css_perm(dptr(X), m, asize, bsize, ll(0), ll(n), vp(st), tres, runs, C.c_uint64(0), vp(states), dist_smem, tracks_smem,
                                 32, 2, wide, chain, dptr(sc), dptr(p), iptr(hits), iptr(nn))
                assert list(hits) == want_h and list(nn) == want_n and list(p) == want_p
            # the general kernel with the large-cohort surrogate (q, digit matrices, labels in global scratch; u8 MMA)
            emu.emu_css_perm_sur.restype = C.c_ulonglong
            for wide, qb in ((1, 23), (0, 9)):
                sc, p, hits, nn = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
                rechecks = emu.emu_css_perm_sur(dptr(X), m, asize, bsize, ll(0), ll(n), vp(st), tres, runs, C.c_uint64(0), vp(states), 0, 0,
                                                32, 2, wide, chain, qb, dptr(sc), dptr(p), iptr(hits), iptr(nn))
                assert list(hits) == want_h and list(nn) == want_n and list(p) == want_p
                if m == 4:
                    assert rechecks > 0                            # the tie window cannot be decided by the surrogate
            sc, p, hits, nn = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
            emu.emu_css_perm2(dptr(X), m, asize, bsize, ll(0), ll(n), vp(st), tres, runs, C.c_uint64(0), vp(states), chain, qbits, 32, 2,
                              dptr(sc), dptr(p), iptr(hits), iptr(nn))
            assert list(hits) == want_h and list(nn) == want_n and list(p) == want_p
            if not chain and 8 <= m <= 64:
                # the headline kernel (two permutations in flight per thread, [word][slot] labels, split LCG, no fp64 distances)
                emu.emu_css_perm3.restype = C.c_ulonglong
                sc3, p3, hits3, nn3 = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
                emu.emu_css_perm3(dptr(X), m, asize, bsize, ll(0), ll(n), vp(st), tres, runs, C.c_uint64(0), vp(states), qbits, 2,
                                  dptr(sc3), dptr(p3), iptr(hits3), iptr(nn3))
                assert list(hits3) == want_h and list(nn3) == want_n and list(p3) == want_p
                assert np.array_equal(sc3, sc)                     # observed score: same bits as the round-1 kernel's


def test_css_perm3_rounds_and_ragged_tails(emu, oracle):
    """fpt_css_perm3_kernel over several rounds of 512 permutations: early stops in the first, second and third round, a ragged
    last round, 20+20 (two k-steps of the u8 MMA) and 9+7 individuals, against the oracle's independent-shuffle restatement"""
    emu.emu_css_perm3.restype = C.c_ulonglong
    rng = np.random.default_rng(18)
    for asize, bsize, n in ((20, 20, 3), (9, 7, 4)):
        m = asize + bsize
        X = rng.normal(size=(n, m, 2))
        X[:, :asize, 0] += np.linspace(0.0, 0.8, n)[:, None]
        st = np.full(n, 2, dtype=np.uint8)
        states = (np.arange(n, dtype=np.uint64) * 104729 + 77)
        qbits = min(21, int(np.floor(np.log2(2 ** 31 / (min(asize, bsize) * m + 1)))))
        for tres, runs in ((1000, 1000), (150, 1300), (400, 1100)):
            want = []
            for w in range(n):
                dist = np.zeros((m, m))
                oracle.fpt_oracle_calc_dist(dptr(X[w].copy()), m, dptr(dist))
                tr = np.arange(m, dtype=np.int32)
                score = oracle.fpt_oracle_css(dptr(dist), m, iptr(tr), iptr(tr[asize:]), asize, bsize)
                h, nn = C.c_int(), C.c_int()
                pv = oracle.fpt_oracle_significance_indep(dptr(dist), m, asize, bsize, score, tres, runs, int(states[w]), C.byref(h), C.byref(nn))
                want.append((score, pv, h.value, nn.value))
            sc, p, hits, nn = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
            emu.emu_css_perm3(dptr(X), m, asize, bsize, ll(0), ll(n), vp(st), tres, runs, C.c_uint64(0), vp(states), qbits, 2,
                              dptr(sc), dptr(p), iptr(hits), iptr(nn))
            assert [(a, b, int(c), int(d)) for a, b, c, d in zip(sc, p, hits, nn)] == want


def test_css_perm_surrogate_many_tiles(emu, oracle):
    """the large-cohort surrogate over several 32-deep k-steps and 8-wide column tiles (m = 70: 3 k-steps, 9 tiles, ragged)"""
    emu.emu_css_perm_sur.restype = C.c_ulonglong
    rng = np.random.default_rng(21)
    asize, bsize, n = 33, 37, 2
    m = asize + bsize
    X = rng.normal(size=(n, m, 2))
    X[1, :asize, 0] += 0.6
    st = np.full(n, 2, dtype=np.uint8)
    states = np.array([5, 77], dtype=np.uint64)
    for chain in (0, 1):
        want = []
        for w in range(n):
            dist = np.zeros((m, m))
            oracle.fpt_oracle_calc_dist(dptr(X[w].copy()), m, dptr(dist))
            tr = np.arange(m, dtype=np.int32)
            score = oracle.fpt_oracle_css(dptr(dist), m, iptr(tr), iptr(tr[asize:]), asize, bsize)
            h, nn = C.c_int(), C.c_int()
            if chain:
                s64 = C.c_uint64(int(states[w]))
                pv = oracle.fpt_oracle_significance(dptr(dist), m, iptr(tr), asize, bsize, score, 1000, 48, C.byref(s64), C.byref(h), C.byref(nn))
            else:
                pv = oracle.fpt_oracle_significance_indep(dptr(dist), m, asize, bsize, score, 1000, 48, int(states[w]), C.byref(h), C.byref(nn))
            want.append((pv, h.value, nn.value))
        sc, p, hits, nn = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        emu.emu_css_perm_sur(dptr(X), m, asize, bsize, ll(0), ll(n), vp(st), 1000, 48, C.c_uint64(0), vp(states), 0, 0, 32, 2, 1, chain, 23,
                             dptr(sc), dptr(p), iptr(hits), iptr(nn))
        assert [(p[w], hits[w], nn[w]) for w in range(n)] == want
        assert 0 < hits[0] < 48                                   # a window where the decisions actually vary


def test_large_cohort_observed_score_shuffle_and_surrogate_distance(emu, oracle):
    """csrc/fpt_css_observed.cuh, the SIMT half of the tensor-memory permutation path: (1) the warp-per-window observed-score
    kernel equals css() of the reference's calc_dist matrix bit for bit; (2) the software-pipelined Fisher-Yates produces the
    labels of fpt_generate_labels (the reference's random_shuffle on its nrand48 stream) and the warp-wide css() of those labels
    equals the oracle's; (3) the Newton/Heron surrogate distance is symmetric bit for bit and within 2^-50 of the exact root."""
    rng = np.random.default_rng(5)
    for asize, bsize in ((37, 30), (5, 64), (1, 3)):
        m, nwin = asize + bsize, 5
        X = rng.normal(size=(nwin, m, 2)) * rng.uniform(0.1, 50.0, size=(nwin, 1, 1))
        if m > 8:
            X[0, 3] = X[0, 5]; X[0, 7] = X[0, 5]; X[0, 1] = X[0, 0]       # coincident points: distance exactly 0 (common in real embeddings)
        status = np.full(nwin, 2, dtype=np.uint8)
        status[3] = 1
        got = np.zeros(nwin)
        emu.emu_css_observed(dptr(X), m, asize, bsize, ll(nwin), vp(status), 2, dptr(got))
        at, bt = np.arange(asize, dtype=np.int32), np.arange(asize, m, dtype=np.int32)
        for w in range(nwin):
            if status[w] != 2:
                assert got[w] == 0.0
                continue
            dist = np.zeros((m, m))
            oracle.fpt_oracle_calc_dist(dptr(X[w].copy()), m, dptr(dist))
            assert got[w] == oracle.fpt_oracle_css(dptr(dist), m, iptr(at), iptr(bt), asize, bsize)
        # shuffles and permuted scores
        nperm = 40
        lf, lr = np.zeros((nperm, m), dtype=np.uint16), np.zeros((nperm, m), dtype=np.uint16)
        sc = np.zeros(nperm)
        Xw = X[0].copy()
        emu.emu_umma_shuffle_and_score(dptr(Xw), m, asize, bsize, C.c_uint64(0x1234ABCD5678), nperm, vp(lf), vp(lr), dptr(sc))
        assert np.array_equal(lf, lr)
        assert all(sorted(row) == list(range(m)) for row in lf.tolist())
        dist = np.zeros((m, m))
        oracle.fpt_oracle_calc_dist(dptr(Xw), m, dptr(dist))
        for k in range(nperm):
            tr = lf[k].astype(np.int32)
            assert sc[k] == oracle.fpt_oracle_css(dptr(dist), m, iptr(tr[:asize].copy()), iptr(tr[asize:].copy()), asize, bsize)
    n = 4000
    pts = [rng.normal(size=n) * 10.0 ** rng.uniform(-3, 3, size=n) for _ in range(4)]
    pts[2][:5], pts[3][:5] = pts[0][:5], pts[1][:5]                  # coincident points: distance 0
    dij, dji = np.zeros(n), np.zeros(n)
    emu.emu_umma_dist(dptr(pts[0]), dptr(pts[1]), dptr(pts[2]), dptr(pts[3]), n, dptr(dij), dptr(dji))
    assert np.array_equal(dij, dji)
    exact = np.hypot(pts[0] - pts[2], pts[1] - pts[3])
    assert np.all(dij[:5] == 0.0)
    rel = np.abs(dij[5:] - exact[5:]) / exact[5:]
    assert rel.max() < 2.0 ** -50


def test_perm3_exact_quotient_magic(emu):
    """fpt_p3_magic: r mod n through one multiply-high and a shift, exact for every 31-bit draw and 2 <= n <= 64 — random
    draws plus the edges where a rounded-up magic number would first fail (multiples of n and their neighbours near 2^31)"""
    emu.emu_p3_magic_check.restype = C.c_longlong
    rng = np.random.default_rng(5)
    rs = [rng.integers(0, 2 ** 31, size=200000, dtype=np.int64)]
    top = 2 ** 31 - 1
    for n in range(2, 65):
        k = top // n
        rs.append(np.array([k * n - 1, k * n, k * n + 1, top, top - 1, (k - 1) * n - 1, (k - 1) * n, n - 1, n, 0, 1], dtype=np.int64))
    rs = np.concatenate(rs).clip(0, top).astype(np.uint32)
    assert emu.emu_p3_magic_check(vp(rs), ll(rs.size)) == 0


def test_large_cohort_shuffle_exact_quotient_magic(emu):
    """fpt_magic31 (the table of the large-cohort shuffle, shift taken from clz(n - 1)): exact remainder and the reference's
    acceptance limit for every n up to 1100, random draws plus the edges near multiples of n and near 2^31"""
    emu.emu_magic31_check.restype = C.c_longlong
    rng = np.random.default_rng(6)
    rs = [rng.integers(0, 2 ** 31, size=4000, dtype=np.int64)]
    top = 2 ** 31 - 1
    for n in (2, 3, 7, 64, 65, 127, 128, 129, 500, 999, 1000, 1023, 1024, 1025, 1100):
        k = top // n
        rs.append(np.array([k * n - 1, k * n, k * n + 1, top, top - 1, (k - 1) * n, n - 1, n, 0, 1], dtype=np.int64))
    rs = np.concatenate(rs).clip(0, top).astype(np.uint32)
    assert emu.emu_magic31_check(vp(rs), ll(rs.size), C.c_uint(1100)) == 0


def test_sturm_counts_by_the_determinant_recurrence(emu):
    """fpt_sturm_count (the eigenvector kernel's multisection probe, also used on the Lanczos tridiagonals up to order 384): the
    number of eigenvalues below x from sign changes of the rescaled determinant recurrence equals numpy's count — random
    matrices, graded ones spanning 12 decades (the rescaling), decoupled blocks (e = 0) and probes ON diagonal entries"""
    rng = np.random.default_rng(8)
    cases = []
    for n in (2, 3, 11, 40, 64, 165, 384):
        d, e = rng.normal(size=n), rng.normal(size=n - 1)
        cases.append((d, e))
        g = 10.0 ** (-12.0 * np.arange(n) / max(1, n - 1))                # graded: entries from 1 down to 1e-12
        cases.append((d * g, e * np.sqrt(g[:-1] * g[1:])))
        e0 = e.copy(); e0[n // 2 - 1 if n > 2 else 0] = 0.0               # two decoupled blocks
        cases.append((d, e0))
        cases.append((np.full(n, 0.3), np.zeros(n - 1)))                  # diagonal matrix, all entries equal
    for d, e in cases:
        n = d.size
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        nrm = np.abs(T).sum(axis=1).max()
        d, e, T = d / nrm, e / nrm, T / nrm
        ev = np.linalg.eigvalsh(T)
        xs = np.concatenate([rng.uniform(-1.1, 1.1, 60), 0.5 * (ev[:-1] + ev[1:]) if n > 1 else [], ev + 1e-7, ev - 1e-7, d[: min(n, 8)]])
        xs = np.ascontiguousarray(xs[np.abs(xs[:, None] - ev[None, :]).min(axis=1) > 1e-9], dtype=np.float64)   # not within rounding of an eigenvalue
        cnt = np.zeros(xs.size, dtype=np.int32)
        e2 = np.ascontiguousarray(np.append(e * e, 0.0))
        emu.emu_sturm_counts(dptr(np.ascontiguousarray(d)), dptr(e2), n, dptr(xs), xs.size, iptr(cnt))
        want = (ev[None, :] < xs[:, None]).sum(axis=1)
        assert np.array_equal(cnt, want), (n, np.flatnonzero(cnt != want)[:5])


def test_fet_log_mode_walk_against_exact_rationals(emu):
    """the score kernel's log-mode walk as compiled for the CPU (term ratios by first differences of numerator and denominator,
    Newton reciprocal, cut-off every fourth term) against exact rational arithmetic of the reference's two-tailed rule, both forms
    of the kernel (plain and tile-sorting)"""
    from math import log10
    import checkers
    rng = np.random.default_rng(23)
    T = []
    for _ in range(1100):
        n1, n2 = int(rng.integers(40, 500)), int(rng.integers(40, 500))
        f = rng.uniform(0.05, 0.95)
        a, c = int(rng.binomial(n1, f)), int(rng.binomial(n2, min(0.99, max(0.01, f + rng.normal(0, 0.08)))))
        T.append((a, n1 - a, c, n2 - c))
    T = np.ascontiguousarray(np.array(T, dtype=np.int32))
    maxn = int(T.sum(1).max())
    plain, srt = np.zeros(len(T)), np.zeros(len(T))
    emu.emu_fet_score(iptr(T), ll(len(T)), maxn, 1, 0, 2, dptr(plain))
    emu.emu_fet_score_sorted(iptr(T), ll(len(T)), maxn, 1, 0, 1, dptr(srt))
    assert np.array_equal(plain, srt)
    for t, g in list(zip(T, plain))[:300]:
        P = checkers.fet_exact_rule(*[int(v) for v in t])
        want = 0.0 if P == 1 else -(log10(P.numerator) - log10(P.denominator))
        assert g == pytest.approx(want, rel=1e-9, abs=1e-10), tuple(t)
