"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes, host buffers), against the CPU
oracle (oracle/liboracle.so) and — where oracle/_ref travelled along — the compiled reference itself, on the
same seeded inputs.

Tolerances (BASELINE.json north_star): contingency tables, window indexing and scored/discarded calls bit-exact;
FET -log10 P within 1e-9 relative; CSS scores within 1e-5 relative; permutation p-values identical for the same
random stream.
"""
import ctypes as C

import numpy as np
import pytest

import checkers
from checkers import dptr, iptr

pytestmark = pytest.mark.gpu

FET_RTOL = 1e-9
CSS_RTOL = 1e-5


def _synth(seed, length, nsnp, asize, bsize, **kw):
    import fpt_b200.synth as synth
    ch = synth.chromosome(seed, length, nsnp, asize, bsize, **kw)
    return ch, synth.reference_layout(ch)


# ------------------------------------------------------------------------------------------------ FET
@pytest.mark.parametrize("asize,bsize,nsnp", [(20, 20, 20000), (11, 10, 3001), (1, 1, 257), (33, 40, 5000), (300, 250, 700)])
def test_fet_per_snp_tables_and_scores(fpt, oracle, asize, bsize, nsnp):
    ch, (av, bv, _, _) = _synth(11 + asize, 10 * nsnp, nsnp, asize, bsize)
    tab_o = np.zeros((nsnp, 4), dtype=np.int32)
    sc_o = np.zeros(nsnp)
    oracle.fpt_oracle_fet_per_snp(dptr(av), dptr(bv), nsnp, asize, bsize, iptr(tab_o), dptr(sc_o))
    for a, b in ((av, bv), (ch["acodes"], ch["bcodes"])):        # reference layout and compact codes
        tab, sc = fpt.fet_per_snp(a, b, asize, bsize)
        assert np.array_equal(tab, tab_o)                        # bit-exact contingency tables
        np.testing.assert_allclose(sc, sc_o, rtol=FET_RTOL, atol=1e-13)
        # P = 1 (score 0) where both say so; a table whose P is 1 up to the rounding of lp0 + log S (|score| < 1e-12) may come out as
        # -0.0 on one side and as 1e-17 on the other (exact mode is operation-identical, the log-mode walk is not)
        assert np.array_equal(np.abs(sc) > 1e-12, np.abs(sc_o) > 1e-12)
        if asize + bsize <= 67:
            assert np.array_equal(sc == 0, sc_o == 0)


def test_fet_exact_domain_all_small_tables(fpt, oracle):
    """every 2x2 table with N <= 24 plus a sample up to N = 67: the exact-arithmetic domain of the reference"""
    t = [(a, b, c, n - a - b - c) for n in range(25) for a in range(n + 1) for b in range(n + 1 - a) for c in range(n + 1 - a - b)]
    rng = np.random.default_rng(5)
    for _ in range(40000):
        n = int(rng.integers(25, 68))
        cuts = np.sort(rng.integers(0, n + 1, size=3))
        t.append((cuts[0], cuts[1] - cuts[0], cuts[2] - cuts[1], n - cuts[2]))
    T = np.array(t, dtype=np.int32)
    so = np.zeros(len(T))
    oracle.fpt_oracle_fet_tables(iptr(T), len(T), dptr(so))
    sg = fpt.fet_tables(T)
    np.testing.assert_allclose(sg, so, rtol=FET_RTOL, atol=1e-13)
    assert np.array_equal(sg == 0, so == 0)
    if checkers.ref_available():                                  # and against the compiled reference's fet()
        rf = checkers.load_ref_fet()
        tmp = (C.c_int * 4)()
        idx = rng.choice(len(T), size=20000, replace=False)
        for i in idx:
            f = (C.c_int * 4)(*[int(v) for v in T[i]])
            pr = rf.fet(f, tmp)
            want = -1.0 * np.log10(pr)
            assert abs(sg[i] - want) <= FET_RTOL * abs(want) + 1e-13


def test_fet_log_mode_high_coverage(fpt, oracle):
    import fpt_b200.synth as synth
    T = synth.coverage_tables(3, 200000, 20, 500)
    so = np.zeros(len(T))
    oracle.fpt_oracle_fet_tables(iptr(T), len(T), dptr(so))
    sg = fpt.fet_tables(T)
    np.testing.assert_allclose(sg, so, rtol=FET_RTOL, atol=1e-12)
    # forcing log mode on small tables agrees with exact mode except on the reference's rounding-dependent ties
    Ts = synth.coverage_tables(4, 50000, 2, 30)
    e, l = fpt.fet_tables(Ts), fpt.fet_tables(Ts, force_log=True)
    close = np.isclose(e, l, rtol=1e-9, atol=1e-12)
    assert close.mean() > 0.99


def test_fet_log_mode_against_exact_rationals(fpt):
    """The GPU's log-space walk (tables beyond the reference's u64 domain, N > 67: SURVEY Q2, fisher/cFisher.c:256-284,473-483) against
    exact rational arithmetic of the reference's two-tailed rule — no oracle in between."""
    from math import log10
    rng = np.random.default_rng(19)
    T = []
    for _ in range(400):
        n1, n2 = int(rng.integers(40, 500)), int(rng.integers(40, 500))
        f = rng.uniform(0.05, 0.95)
        a, c = int(rng.binomial(n1, f)), int(rng.binomial(n2, min(0.99, max(0.01, f + rng.normal(0, 0.08)))))
        T.append((a, n1 - a, c, n2 - c))
    T = np.ascontiguousarray(np.array(T, dtype=np.int32))
    assert (T.sum(axis=1) > 67).all()
    got = fpt.fet_tables(T)
    for t, g in zip(T, got):
        P = checkers.fet_exact_rule(*[int(v) for v in t])
        want = 0.0 if P == 1 else -(log10(P.numerator) - log10(P.denominator))
        assert g == pytest.approx(want, rel=1e-9, abs=1e-10), tuple(t)


@pytest.mark.parametrize("semantics", [0, 1])
@pytest.mark.parametrize("geom", [(2500, 500), (1000, 1000), (700, 300)])
def test_fet_scan_matches_oracle(fpt, oracle, semantics, geom):
    wsize, wstep = geom
    regend, nsnp, asize, bsize, seed = 150000, 4000, 20, 20, 99
    ch, (av, bv, apos, bpos) = _synth(21, regend, nsnp, asize, bsize)
    n = regend // wstep
    s_o, d_o = np.zeros(n), np.zeros(n)
    assert oracle.fpt_oracle_fet_scan(dptr(av), dptr(bv), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, av.size, bv.size,
                                      0.95, dptr(s_o), dptr(d_o), semantics, seed) == 0
    s_g, d_g, wr = fpt.fet_scan(av, bv, ch["pos"], asize, bsize, regend, wsize, wstep, 0.95, semantics=semantics, seed=seed)
    np.testing.assert_allclose(s_g, s_o, rtol=FET_RTOL, atol=1e-13)
    np.testing.assert_allclose(d_g, d_o, rtol=1e-9, atol=1e-13)
    # sharded: two window ranges reproduce the full scan exactly (streams are keyed by global window index)
    h = n // 2 + 3
    s1, d1, _ = fpt.fet_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 0.95, semantics=semantics,
                             seed=seed, window_begin=0, window_end=h)
    s2, d2, _ = fpt.fet_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 0.95, semantics=semantics,
                             seed=seed, window_begin=h, window_end=n)
    assert np.array_equal(np.concatenate([s1, s2]), s_g) and np.array_equal(np.concatenate([d1, d2]), d_g)


def test_fet_dropin_matches_reference(fpt, ref_fet):
    """the drop-in Cython-signature call against the reference's own serial `compute` (deterministic column)"""
    import fpt_b200.fisher_cython as serial
    import fpt_b200.fisher_cython_parallel as par
    regend, wsize, wstep, nsnp, asize, bsize = 120000, 2500, 500, 3000, 20, 20
    ch, (av, bv, apos, bpos) = _synth(31, regend, nsnp, asize, bsize)
    n = regend // wstep
    s_r, d_r = np.zeros(n), np.zeros(n)
    with checkers.silence_stdout():
        ref_fet.compute(dptr(av), dptr(bv), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, av.size, bv.size, 0.95, dptr(s_r), dptr(d_r))
    for mod in (serial, par):
        s_g, d_g = np.zeros(n), np.zeros(n)
        mod.fisher_exact_tester(av, bv, apos, bpos, 0, regend, wsize, wstep, av.size, bv.size, 0.95, s_g, d_g)
        np.testing.assert_allclose(s_g, s_r, rtol=FET_RTOL, atol=1e-13)
    # threaded semantics: nothing is computed below 103 windows
    s_g, d_g = np.zeros(100), np.zeros(100)
    par.fisher_exact_tester(av, bv, apos, bpos, 0, 50000, wsize, wstep, av.size, bv.size, 0.95, s_g, d_g)
    assert not s_g.any() and not d_g.any()


def test_fet_window_bootstrap_matches_reference_stream(fpt, ref_fet):
    """one window, explicit 48-bit state: sigma equals the reference's fisher_exact_test run from that state"""
    asize = bsize = 20
    for npos, state in ((37, 0x1234ABCD5678), (5, 42), (1, 7), (300, 0xFFFFFFFFFFFF), (1000, 99)):
        ch, (av, bv, _, _) = _synth(npos, 100000, npos, asize, bsize)
        res = np.zeros(2)
        f, tmp = (C.c_int * 4)(), (C.c_int * 4)()
        samples, stds, fets = np.zeros(npos), np.zeros(100), np.zeros(npos)
        st = checkers.state_to_ushort3(state)
        ref_fet.fisher_exact_test(dptr(res), dptr(av), dptr(bv), asize, bsize, npos, f, tmp, dptr(samples), dptr(stds), 100, dptr(fets), st, 0.95)
        pos = np.arange(npos, dtype=np.int32)
        s, d, w = fpt.fet_scan(av, bv, pos, asize, bsize, 100000, 100000, 100000, 0.95, states=np.array([state], dtype=np.uint64))
        assert w[0] == 1
        if npos > 1:                                   # n == 1 makes the reference read one past its array (Q8)
            np.testing.assert_allclose(s[0], res[0], rtol=FET_RTOL, atol=1e-13)
            np.testing.assert_allclose(d[0], res[1], rtol=1e-9, atol=1e-13)


# ------------------------------------------------------------------------------------------------ CSS
def _css_oracle_scan(oracle, av, bv, apos, bpos, regend, wsize, wstep, mct, mcr, mds, semantics, seed, dros=0):
    n = regend // wstep
    s_o, p_o = np.zeros(n), np.zeros(n)
    assert oracle.fpt_oracle_css_scan(dptr(av), dptr(bv), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, av.size, bv.size,
                                      mct, mcr, dros, mds, dptr(s_o), dptr(p_o), semantics, seed) == 0
    return s_o, p_o


def _oracle_window_delta(oracle, ch, av, bv, asize, bsize, w, wsize, wstep):
    """filled dissimilarity matrix of window w (compare_all + fill_averages restated by the oracle)"""
    l, r = C.c_int64(0), C.c_int64(0)
    oracle.fpt_oracle_window_bounds(iptr(ch["pos"]), ch["pos"].size, w, wsize, wstep, C.byref(l), C.byref(r))
    l, r = l.value, r.value
    m = asize + bsize
    D = np.zeros((m, m))
    oracle.fpt_oracle_compare_all(dptr(av[l * asize:r * asize].copy()), dptr(bv[l * bsize:r * bsize].copy()), asize, bsize, r - l, dptr(D))
    assert oracle.fpt_oracle_fill_averages(dptr(D), m)
    return D


@pytest.mark.parametrize("mds", [1, 2])
@pytest.mark.parametrize("shape", [(20, 20, 2500, 500), (7, 9, 3000, 1000), (70, 60, 6000, 3000)])
def test_css_smacof_stage_matches_oracle_bit_for_bit(fpt, oracle, mds, shape):
    """SMACOF (css.c:852-938) from a given start is the reference's arithmetic operation for operation: iteration count, final
    stress and final coordinates of every (window, start) equal the oracle's bit for bit — including the windows whose
    stopping decision `sigma_prev - sigma > 1e-6` falls within rounding of the threshold, where the stress is re-summed in
    the reference's order (css.c:767-777). mds 1: the four drand48 starts of the window's stream; mds 2: the start is this
    library's own classical-MDS embedding (probe of an mds 0 scan), handed to the oracle's smacof."""
    asize, bsize, wsize, wstep = shape
    m = asize + bsize
    regend, nsnp, seed = 60000, 1500, 5
    ch, (av, bv, apos, bpos) = _synth(41 + asize, regend, nsnp, asize, bsize)
    s_g, p_g, wr, pr = fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 10, 200, mds=mds,
                                    seed=seed, probes=True)
    X0 = None
    if mds == 2:
        X0 = fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 0, 0, mds=0, seed=seed, probes=True)[3]["X"]
    nruns = 4 if mds == 1 else 1
    checked = 0
    for w in np.flatnonzero(wr == 1):
        D = _oracle_window_delta(oracle, ch, av, bv, asize, bsize, w, wsize, wstep)
        state = C.c_uint64(oracle.fpt_oracle_window_state(seed, int(w), 1))
        best, best_X = None, None
        for run in range(nruns):
            if mds == 1:
                Xs = np.array([oracle.fpt_oracle_drand48(C.byref(state)) for _ in range(2 * m)]).reshape(m, 2)
            else:
                Xs = X0[w].copy()
            if not np.isfinite(Xs).all():
                break
            k = C.c_int(0)
            sig = oracle.fpt_oracle_smacof(dptr(D), m, dptr(Xs), 300, 1e-6, C.byref(k))
            assert pr["smacof_iters"][w, run] == k.value, "window %d start %d" % (w, run)
            assert pr["smacof_sigma"][w, run] == sig
            if best is None or sig < best:
                best, best_X = sig, Xs
            checked += 1
        else:
            assert np.array_equal(pr["X"][w], best_X)
    assert checked >= 10 * nruns


@pytest.mark.parametrize("mds", [0, 1, 2])
@pytest.mark.parametrize("shape", [(20, 20, 2500, 500), (7, 9, 3000, 1000), (2, 2, 2500, 500)])
def test_css_scan_matches_oracle(fpt, oracle, mds, shape):
    asize, bsize, wsize, wstep = shape
    regend, nsnp, seed = 60000, 1500, 5
    ch, (av, bv, apos, bpos) = _synth(41 + asize, regend, nsnp, asize, bsize)
    s_o, p_o = _css_oracle_scan(oracle, av, bv, apos, bpos, regend, wsize, wstep, 10, 200, mds, 0, seed)
    s_g, p_g, wr, pr = fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 10, 200, mds=mds,
                                    seed=seed, probes=True)
    # same windows scored / discarded — except where the score is the reference's own "discarded" sentinel -1.0
    # (css.c:126: `if (result != -1)`), which a 2+2 toy cohort can hit exactly and where the last bit decides
    near_sentinel = (pr["status"] == 2) & (np.abs(np.where(wr == 1, s_g, -1.0) + 1.0) < 1e-9)
    assert np.array_equal((wr == 1) | near_sentinel, (p_o != 0) | near_sentinel)
    finite = np.isfinite(s_o) & np.isfinite(s_g) & ~near_sentinel
    if mds != 1:
        # with lambda2 ~ lambda3 the embedding depends on the eigensolver's arbitrary basis
        ev = pr["evals"]
        finite &= ~((wr == 1) & ((ev[:, 1] - ev[:, 2]) < 1e-8 * np.maximum(ev[:, 0], 1e-300)))
    assert finite.sum() >= (0.8 if asize + bsize > 4 else 0.4) * (wr == 1).sum()
    rel = np.abs(s_g[finite] - s_o[finite]) / np.maximum(np.abs(s_o[finite]), 1e-300)
    bad = rel > CSS_RTOL
    if mds == 2:
        # mds 2 starts SMACOF from the classical-MDS embedding, i.e. from the eigensolver's output (GSL in the reference,
        # tred2/tql2 in the oracle, Sturm + inverse iteration here: equal to ~1e-14, not bit for bit — DESIGN.md "GSL boundary").
        # SMACOF itself is pinned bit for bit from a given start (test_css_smacof_stage_matches_oracle_bit_for_bit); a start
        # that differs in its last bits changes the iteration count only where the reference's own stopping rule
        # `sigma_prev - sigma > 1e-6` is decided within 1 % of the threshold — those windows depend on the GSL build in the
        # reference too. They are identified from the ORACLE's margin and its iteration count, nothing else is excused.
        n = regend // wstep
        degenerate = np.zeros(n, dtype=bool)
        for w in np.flatnonzero(wr == 1):
            D = _oracle_window_delta(oracle, ch, av, bv, asize, bsize, w, wsize, wstep)
            Xo, evo = np.zeros((asize + bsize, 2)), np.zeros(3)
            oracle.fpt_oracle_cmds(dptr(D), asize + bsize, dptr(Xo), dptr(evo))
            k, margin = C.c_int(0), C.c_double(0)
            oracle.fpt_oracle_smacof_margin(dptr(D), asize + bsize, dptr(Xo), 300, 1e-6, C.byref(k), C.byref(margin))
            if k.value != pr["smacof_iters"][w, 0]:
                assert margin.value < 1e-8, "window %d: %d vs %d iterations at margin %g" % (w, pr["smacof_iters"][w, 0], k.value, margin.value)
                degenerate[w] = True
        assert degenerate.sum() <= max(1, int(0.02 * finite.sum()))
        bad &= ~degenerate[finite]
    assert bad.sum() == 0, "CSS score mismatches: %d (max rel %g)" % (bad.sum(), rel.max())
    agree = ~bad
    if asize + bsize > 4:
        assert np.array_equal(p_g[finite][agree], p_o[finite][agree])  # identical permutation p-values
    # (with 2+2 individuals most permutations tie with the observed score exactly, so `>=` is decided by the last
    #  bit of the embedding, which differs between eigensolvers — also between GSL and any other; the p-value
    #  arithmetic itself is pinned bit-for-bit by test_css_significance_matches_reference_stream)
    # f64 reference layout gives the same bits as compact codes
    s_f, p_f, _ = fpt.css_scan(av, bv, ch["pos"], asize, bsize, regend, wsize, wstep, 10, 200, mds=mds, seed=seed)
    assert np.array_equal(s_f, s_g, equal_nan=True) and np.array_equal(p_f, p_g)


@pytest.mark.parametrize("mct,mcr", [(1000, 1000), (10, 5000), (1, 300), (0, 100), (3, 0)])
def test_css_permutation_early_stop_matches_oracle(fpt, oracle, mct, mcr):
    asize, bsize, wsize, wstep, regend, nsnp, seed = 20, 20, 2500, 500, 40000, 1000, 17
    ch, (av, bv, apos, bpos) = _synth(61, regend, nsnp, asize, bsize, planted_every=4, planted_len=10)
    s_o, p_o = _css_oracle_scan(oracle, av, bv, apos, bpos, regend, wsize, wstep, mct, mcr, 0, 0, seed)
    s_g, p_g, wr = fpt.css_scan(av, bv, ch["pos"], asize, bsize, regend, wsize, wstep, mct, mcr, mds=0, seed=seed)
    np.testing.assert_allclose(s_g, s_o, rtol=CSS_RTOL, atol=1e-12)
    assert np.array_equal(p_g, p_o)


@pytest.mark.parametrize("asize,bsize,mct,mcr", [(20, 20, 1000, 1000), (20, 20, 10, 5000), (7, 9, 300, 1300), (30, 34, 50, 600), (4, 4, 5, 100), (13, 20, 40, 700)])
def test_css_headline_permutation_kernel_matches_round1_kernel(fpt, oracle, asize, bsize, mct, mcr):
    """cohorts of 8..64: fpt_css_perm3_kernel (default) and the round-1 kernel fpt_css_perm2_kernel take the same decisions —
    scores bit-identical, hits, permutations drawn and p identical — over several rounds, early stops and ragged tails; and
    both equal the oracle (css.c:727-752 restated)"""
    from fpt_b200 import api
    regend, wsize, wstep, nsnp, seed = 60000, 2500, 500, 1500, 23
    ch, (av, bv, apos, bpos) = _synth(61 + asize, regend, nsnp, asize, bsize, planted_every=4, planted_len=10)
    out = []
    try:
        for v in (1, 0):
            api.set_perm_small_kernel(v)
            out.append(fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, mct, mcr, mds=0, seed=seed, probes=True))
    finally:
        api.set_perm_small_kernel(1)
    (s3, p3, w3, pr3), (s2, p2, w2, pr2) = out
    assert (w3 == 1).sum() > 50 and np.array_equal(w3, w2)
    assert np.array_equal(s3, s2, equal_nan=True) and np.array_equal(p3, p2)
    assert np.array_equal(pr3["hits"], pr2["hits"]) and np.array_equal(pr3["nperm"], pr2["nperm"])
    s_o, p_o = _css_oracle_scan(oracle, av, bv, apos, bpos, regend, wsize, wstep, mct, mcr, 0, 0, seed)
    np.testing.assert_allclose(s3, s_o, rtol=CSS_RTOL, atol=1e-12)
    if asize + bsize > 8:
        # 4+4: one permutation in 70 swaps the two groups as sets and ties with the observed score up to rounding, so `>=` is
        # decided by the last bits of the embedding, i.e. by the eigensolver (GSL boundary, see test_css_scan_matches_oracle)
        assert np.array_equal(p3, p_o)


@pytest.mark.parametrize("asize,bsize", [(20, 20), (1, 2), (16, 16), (13, 20), (25, 23), (7, 9)])
def test_css_register_tridiagonalisation_matches_shared_memory_kernel(fpt, oracle, asize, bsize):
    """cohorts of 3..48, classical MDS: the tridiagonalisation with the matrix in registers (default, csrc/fpt_css_eig_reg.cuh,
    paddings 32 / 40 / 48) and the shared-memory kernel (csrc/fpt_css_eig.cuh) keep the same windows and agree to rounding on
    scores and eigenvalues; both are within tolerance of the oracle's cmds + css (css.c:505-560,608-647)"""
    from fpt_b200 import api
    regend, wsize, wstep, nsnp, seed = 60000, 2500, 500, 1800, 29
    ch, (av, bv, apos, bpos) = _synth(71 + asize, regend, nsnp, asize, bsize, planted_every=4, planted_len=10)
    out = []
    try:
        for v in (1, 0):
            api.set_mds_small_kernel(v)
            out.append(fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 50, 200, mds=0, seed=seed, probes=True))
    finally:
        api.set_mds_small_kernel(1)
    (s1, p1, w1, pr1), (s0, p0, w0, pr0) = out
    assert np.array_equal(w1, w0) and (w1 == 1).sum() > 40
    fin = np.isfinite(s0) & (w0 == 1)
    assert np.array_equal(np.isfinite(s1), np.isfinite(s0))
    gap_ok = (pr0["evals"][:, 1] - pr0["evals"][:, 2]) > 1e-6 * np.abs(pr0["evals"][:, 0])       # lambda2 ~ lambda3: any solver's choice
    sel = fin & gap_ok
    assert sel.sum() > 30
    np.testing.assert_allclose(s1[sel], s0[sel], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(pr1["evals"][sel, :2], pr0["evals"][sel, :2], rtol=1e-10, atol=1e-10 * np.abs(pr0["evals"][sel, 0]).max())
    s_o, p_o = _css_oracle_scan(oracle, av, bv, apos, bpos, regend, wsize, wstep, 50, 200, 0, 0, seed)
    np.testing.assert_allclose(s1[sel], s_o[sel], rtol=CSS_RTOL, atol=1e-12)
    if asize + bsize > 4:
        assert np.array_equal(p1[sel], p_o[sel])


def test_css_dropin_matches_reference(fpt, ref_css):
    """drop-in call vs the reference's serial `compute`, classical MDS (deterministic score column)"""
    import fpt_b200.css_cython as serial
    regend, wsize, wstep, nsnp, asize, bsize = 60000, 2500, 500, 1500, 20, 20
    ch, (av, bv, apos, bpos) = _synth(71, regend, nsnp, asize, bsize)
    n = regend // wstep
    s_r, p_r = np.zeros(n), np.zeros(n)
    with checkers.silence_stdout():
        ref_css.compute(dptr(av), dptr(bv), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, av.size, bv.size, 10, 50, 0, 0, dptr(s_r), dptr(p_r))
    s_g, p_g = np.zeros(n), np.zeros(n)
    serial.cluster_separation_scorer(av, bv, apos, bpos, 0, regend, wsize, wstep, av.size, bv.size, 10, 50, 0, 0, s_g, p_g)
    np.testing.assert_allclose(s_g, s_r, rtol=CSS_RTOL, atol=1e-12)
    assert np.array_equal(p_g != 0, p_r != 0)


def test_css_significance_matches_reference_stream(fpt, ref_css, oracle):
    """chain mode, per window, explicit state: p equals the reference's significance_treshold run on fresh identity labels"""
    asize, bsize, wsize, wstep, regend, nsnp = 12, 9, 5000, 5000, 40000, 900
    m = asize + bsize
    ch, (av, bv, apos, bpos) = _synth(81, regend, nsnp, asize, bsize)
    n = regend // wstep
    states = np.array([0x5EED0000 + 977 * i for i in range(n)], dtype=np.uint64)
    fpt.set_perm_mode(True)
    try:
        s_g, p_g, wr, pr = fpt.css_scan(av, bv, ch["pos"], asize, bsize, regend, wsize, wstep, 5, 1400, mds=0, states_perm=states, probes=True)
    finally:
        fpt.set_perm_mode(False)
    checked = 0
    for w in range(n):
        if not wr[w]:
            continue
        X = checkers.RefMatrix(m, 2, pr["X"][w])
        dist = checkers.RefMatrix(m, m)
        ref_css.calc_dist(X.pp, dist.pp, m)
        tracks = np.arange(m, dtype=np.int32)
        score = ref_css.css(dist.pp, iptr(tracks), iptr(tracks[asize:]), asize, bsize)
        assert score == s_g[w]                       # same embedding -> bit-identical score arithmetic
        st = checkers.state_to_ushort3(int(states[w]))
        p_ref = ref_css.significance_treshold(dist.pp, iptr(tracks), asize, bsize, score, 5, 1400, st)
        assert p_ref == p_g[w]
        checked += 1
    assert checked >= 4


def test_css_independent_shuffles_match_reference_functions(fpt, ref_css, oracle):
    """default mode: permutation k = the reference's random_shuffle on fresh identity labels, stream k*(m-1) draws in;
    hits / permutations drawn / p recomputed with the reference's own random_shuffle + css, per window"""
    asize, bsize, wsize, wstep, regend, nsnp = 20, 20, 5000, 5000, 50000, 1200
    m = asize + bsize
    ch, (av, bv, apos, bpos) = _synth(83, regend, nsnp, asize, bsize)
    n = regend // wstep
    states = np.array([0xC0FFEE00 + 7919 * i for i in range(n)], dtype=np.uint64)
    tres, runs = 6, 600
    s_g, p_g, wr, pr = fpt.css_scan(av, bv, ch["pos"], asize, bsize, regend, wsize, wstep, tres, runs, mds=0, states_perm=states, probes=True)
    checked = 0
    for w in range(n):
        if not wr[w]:
            continue
        X = checkers.RefMatrix(m, 2, pr["X"][w])
        dist = checkers.RefMatrix(m, m)
        ref_css.calc_dist(X.pp, dist.pp, m)
        ident = np.arange(m, dtype=np.int32)
        score = ref_css.css(dist.pp, iptr(ident), iptr(ident[asize:]), asize, bsize)
        assert score == s_g[w]
        hits = k = 0
        while hits < tres and k < runs:
            st = checkers.state_to_ushort3(oracle.fpt_oracle_lcg_skip(int(states[w]), k * (m - 1)))
            tr = np.arange(m, dtype=np.int32)
            ref_css.random_shuffle(iptr(tr), m, st)
            hits += ref_css.css(dist.pp, iptr(tr), iptr(tr[asize:]), asize, bsize) >= score
            k += 1
        assert (hits, k) == (pr["hits"][w], pr["nperm"][w])
        assert p_g[w] == (hits + 1) * 1.0 / (k + 1)
        checked += 1
    assert checked >= 4


@pytest.mark.parametrize("asize,bsize", [(36, 36), (70, 60), (150, 140)])
def test_css_larger_cohorts_take_the_fallback_paths(fpt, oracle, asize, bsize):
    """72: one warp per CTA for classical MDS, gather surrogate; 130: gather surrogate, bigger tiles; 290: Lanczos classical
    MDS and the general permutation kernel (16-bit labels, global-memory scratch, tensor-core surrogate)"""
    regend, wsize, wstep, nsnp, seed = 12000, 3000, 1500, 500, 3
    ch, (av, bv, apos, bpos) = _synth(100 + asize, regend, nsnp, asize, bsize)
    s_o, p_o = _css_oracle_scan(oracle, av, bv, apos, bpos, regend, wsize, wstep, 5, 60, 0, 0, seed)
    s_g, p_g, wr, pr = fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 5, 60, mds=0, seed=seed, probes=True)
    assert np.array_equal(wr == 1, p_o != 0) and (wr == 1).sum() >= 4
    np.testing.assert_allclose(s_g, s_o, rtol=CSS_RTOL, atol=1e-12)
    assert np.array_equal(p_g, p_o)


def test_css_large_cohort_500_plus_500(fpt, oracle):
    """BASELINE configs[4] cohort size (500+500 individuals, 50 kb windows) on a handful of windows, one ragged batch of 40
    permutations: the default large-cohort route (Lanczos classical MDS, observed scores by a warp per window, tcgen05 /
    tensor-memory permutation kernel); the 1000-permutation case is test_css_large_cohort_default_route_1000_permutations_vs_oracle"""
    asize = bsize = 500
    regend, wsize, wstep, nsnp, seed = 150000, 50000, 50000, 500, 4
    ch, (av, bv, apos, bpos) = _synth(500, regend, nsnp, asize, bsize)
    s_o, p_o = _css_oracle_scan(oracle, av, bv, apos, bpos, regend, wsize, wstep, 3, 40, 0, 0, seed)
    s_g, p_g, wr = fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 3, 40, mds=0, seed=seed)
    assert np.array_equal(wr == 1, p_o != 0) and (wr == 1).sum() == 3
    np.testing.assert_allclose(s_g, s_o, rtol=CSS_RTOL, atol=1e-12)
    assert np.array_equal(p_g, p_o)


@pytest.mark.parametrize("asize,bsize,mct,mcr,seed", [(500, 500, 1000, 1000, 4), (500, 500, 150, 1000, 4), (512, 512, 150, 1000, 5),
                                                      (503, 499, 150, 1000, 6)])
def test_css_large_cohort_default_route_1000_permutations_vs_oracle(fpt, oracle, asize, bsize, mct, mcr, seed):
    """BASELINE configs[4] as it is benchmarked: the DEFAULT large-cohort route (Lanczos classical MDS ->
    fpt_css_observed_kernel -> the tcgen05 / tensor-memory permutation kernel fpt_css_perm_umma_kernel) at
    mcT / mcR = 1000, i.e. eight batches of 128 permutations per window (accumulator double-buffer phase flips, ring
    wrap-around across batches, ndone / early-stop bookkeeping), against the CPU oracle (css.c:727-752 restated):
    full 1000 permutations; early stops that land in batch 2..5; m = 1024 (no K / N padding); m = 1002 with unequal
    groups (not a multiple of 4 or 8: Lanczos form 0, padded K / N tiles)."""
    regend, wsize, wstep, nsnp = 150000, 50000, 50000, 500
    ch, (av, bv, apos, bpos) = _synth(500 + asize, regend, nsnp, asize, bsize)
    s_o, p_o = _css_oracle_scan(oracle, av, bv, apos, bpos, regend, wsize, wstep, mct, mcr, 0, 0, seed)
    s_g, p_g, wr, pr = fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, mct, mcr, mds=0,
                                    seed=seed, probes=True)
    assert np.array_equal(wr == 1, p_o != 0) and (wr == 1).sum() == 3
    np.testing.assert_allclose(s_g, s_o, rtol=CSS_RTOL, atol=1e-12)
    assert np.array_equal(p_g, p_o)                                   # identical hits / permutations drawn
    assert np.array_equal(p_g, (pr["hits"] + 1.0) / (pr["nperm"] + 1.0))
    if mct < mcr:                                                     # an early stop beyond the second batch of 128
        stopped = pr["nperm"][(pr["hits"] == mct)]
        assert stopped.size and stopped.max() > 256 and stopped.max() < mcr
    else:
        assert np.all(pr["nperm"] == mcr)


@pytest.mark.parametrize("asize,bsize,nsnp,wsize", [(500, 500, 600, 50000), (130, 171, 700, 50000), (200, 184, 2400, 50000), (150, 150, 6000, 25000),
                                                    (77, 90, 9000, 100000)])
def test_css_genotype_gemm_counts_equal_a_host_product(fpt, asize, bsize, nsnp, wsize):
    """the tcgen05 u8 GEMM D = [P|M][M|P]' (and the popcount kernel) against numpy's integer product of the same indicator
    matrices, every pair of every probed window: exact. Covers tile padding (m = 301, 167), one- and two-byte codes, K chunks
    (192 SNPs each) with accumulation in tensor memory including ragged last chunks, window edges inside a bit-plane word."""
    regend = 200000
    ch, _ = _synth(900 + nsnp, regend, nsnp, asize, bsize)
    m = asize + bsize
    pos = ch["pos"]
    G = np.concatenate([ch["acodes"].reshape(-1, asize), ch["bcodes"].reshape(-1, bsize)], axis=1)
    for w in (0, 1, regend // wsize - 1):
        l, r = int(np.searchsorted(pos, w * wsize, "left")), int(np.searchsorted(pos, w * wsize + wsize, "right"))
        P, M = (G[l:r] == 3).astype(np.int64), (G[l:r] == -3).astype(np.int64)
        want = P.T @ M + M.T @ P
        for mode in (2, 1):
            got = fpt.k4_counts(ch["acodes"], ch["bcodes"], pos, asize, bsize, regend, wsize, wsize, w, mode=mode)
            assert np.array_equal(got, want), "window %d mode %d: %d of %d pairs differ" % (w, mode, int((got != want).sum()), m * m)


@pytest.mark.parametrize("asize,bsize,nsnp,wsize", [(500, 500, 600, 50000), (130, 171, 700, 50000), (200, 184, 2400, 50000), (150, 150, 6000, 25000)])
def test_css_genotype_gemm_on_tensor_cores_matches_popcounts(fpt, asize, bsize, nsnp, wsize):
    """Large cohorts: the genotype-distance matrix D = [P|M][M|P]' (compare_all, css.c:277-327) as ONE u8 GEMM per window on
    tcgen05 / tensor memory (fpt_css_k4_umma_kernel, the default) against the bit-plane popcount kernel writing the same count
    codes: integers, so embedding, scores and p-values must be IDENTICAL — m = 1000; m = 301 (row / column padding of the 128 x 256
    tiles); ~600 SNPs per window (two-byte codes, four K chunks with accumulation in tensor memory); ~750 SNPs per window with
    ragged last chunks. The round-1 route (fp64 matrix, mode 0) agrees to rounding."""
    from fpt_b200 import api
    regend, wstep, seed = 200000, wsize, 21
    ch, _ = _synth(900 + nsnp, regend, nsnp, asize, bsize)
    out = []
    try:
        for mode in (2, 1, 0):
            api.set_k4_mode(mode)
            out.append(fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 20, 100, mds=0, seed=seed, probes=True))
    finally:
        api.set_k4_mode(2)
    (s2, p2, w2, pr2), (s1, p1, w1, pr1), (s0, p0, w0, pr0) = out
    assert (w2 == 1).sum() >= 3 and np.array_equal(w2, w1) and np.array_equal(w2, w0)
    assert np.array_equal(pr2["X"], pr1["X"], equal_nan=True)           # same integers in, same arithmetic after
    assert np.array_equal(s2, s1, equal_nan=True) and np.array_equal(p2, p1)
    np.testing.assert_allclose(s0, s2, rtol=1e-9, atol=1e-12)
    assert np.array_equal(p0, p2)
    np.testing.assert_allclose(pr0["evals"], pr2["evals"], rtol=1e-10, atol=1e-9)


def test_tcgen05_plumbing_against_a_host_product():
    """csrc/fpt_umma.cuh on its own: bulk TMA copies into shared memory, shared-memory descriptors of the K-major core-matrix
    layout, `tcgen05.mma kind::i8` into tensor memory, `tcgen05.ld` back — a 128 x 256 x 128 u8 product equal to the host's."""
    import os
    import subprocess
    probe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "probes", "umma_probe")
    if not os.path.exists(probe):
        pytest.skip("tests/probes/umma_probe not built (__graft_entry__.build() compiles it)")
    r = subprocess.run([probe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "UMMA_PROBE OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("asize,bsize,mct,mcr", [(500, 500, 300, 300), (400, 300, 7, 400), (130, 170, 1000, 1000), (513, 511, 140, 140)])
def test_css_tensor_memory_permutation_kernel_matches_general_kernel(fpt, asize, bsize, mct, mcr):
    """Large cohorts: the tcgen05 / tensor-memory permutation kernel (csrc/fpt_css_perm_umma.cuh) takes the same decisions as
    the general kernel (csrc/fpt_css_perm_large.cuh) — hits, permutations drawn, p and score identical — over several batches
    of 128 permutations, ragged last batches, unequal groups, early stops and K / N padding (m = 300, 700, 1000, 1024)."""
    from fpt_b200 import api
    regend, wsize, wstep, nsnp, seed = 200000, 50000, 50000, 600, 11
    ch, _ = _synth(300 + asize, regend, nsnp, asize, bsize)
    out = []
    try:
        for tm in (1, 0, 2):
            api.set_perm_large_kernel(tm)
            fpt.css_perm_rechecks()
            out.append(fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, mct, mcr, mds=0, seed=seed, probes=True))
            rechecks = fpt.css_perm_rechecks()
    finally:
        api.set_perm_large_kernel(1)
    (s1, p1, w1, pr1), (s0, p0, w0, pr0), (s2, p2, w2, pr2) = out
    # mode 2 (10-bit surrogate) decides a large share of the permutations by the warp-wide exact re-scoring: same answers
    assert rechecks > 10
    assert np.array_equal(s2, s1, equal_nan=True) and np.array_equal(p2, p1)
    assert np.array_equal(pr2["hits"], pr1["hits"]) and np.array_equal(pr2["nperm"], pr1["nperm"])
    assert (w1 == 1).sum() == 4 and np.array_equal(w1, w0)
    assert np.array_equal(s1, s0, equal_nan=True) and np.array_equal(p1, p0)
    assert np.array_equal(pr1["hits"], pr0["hits"]) and np.array_equal(pr1["nperm"], pr0["nperm"])
    assert pr1["nperm"][w1 == 1].max() > 0


@pytest.mark.parametrize("mds", [0, 2])
def test_css_cohort_beyond_the_warp_path(fpt, oracle, mds):
    """150+150 individuals: classical MDS by the Lanczos kernel (CTA per window, matrices in global scratch), permutations by
    the general kernel with 16-bit labels and the tensor-core surrogate"""
    asize = bsize = 150
    regend, wsize, wstep, nsnp, seed = 400000, 20000, 10000, 4000, 9
    ch, (av, bv, apos, bpos) = _synth(77, regend, nsnp, asize, bsize, wstep=wstep)
    s_o, p_o = _css_oracle_scan(oracle, av, bv, apos, bpos, regend, wsize, wstep, 5, 60, mds, 0, seed)
    s_g, p_g, wr = fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 5, 60, mds=mds, seed=seed)
    assert np.array_equal(wr == 1, p_o != 0) and (wr == 1).sum() >= 30
    np.testing.assert_allclose(s_g, s_o, rtol=CSS_RTOL, atol=1e-12)
    assert np.array_equal(p_g, p_o)


def test_css_frequency_metric(fpt, oracle):
    """drosophila = 1: two frequency tracks, m = 2 (compare_freq)"""
    rng = np.random.default_rng(3)
    nsnp, regend, wsize, wstep = 2000, 100000, 5000, 2500
    pos = np.sort(rng.choice(regend, size=nsnp, replace=False)).astype(np.int32)
    fa, fb = rng.random(nsnp), rng.random(nsnp)
    s_o, p_o = _css_oracle_scan(oracle, fa, fb, pos, pos, regend, wsize, wstep, 5, 20, 0, 0, 1, dros=1)
    s_g, p_g, wr = fpt.css_scan(fa, fb, pos, 1, 1, regend, wsize, wstep, 5, 20, drosophila=1, mds=0, seed=1)
    ok = np.isfinite(s_o) & np.isfinite(s_g)
    assert ok.sum() > 0
    np.testing.assert_allclose(s_g[ok], s_o[ok], rtol=CSS_RTOL, atol=1e-12)


def test_position_mismatch_is_an_error(fpt):
    import fpt_b200.fisher_cython as serial
    ch, (av, bv, apos, bpos) = _synth(91, 50000, 500, 4, 4)
    bpos = bpos.copy()
    bpos[40:44] += 1
    s, d = np.zeros(100), np.zeros(100)
    with pytest.raises(fpt.FptError) as e:
        serial.fisher_exact_tester(av, bv, apos, bpos, 0, 50000, 2500, 500, av.size, bv.size, 0.95, s, d)
    assert e.value.code == -3
    # the A/B check runs beside the scan (off the critical path): a failed call leaves the caller's pre-zeroed outputs untouched
    assert not s.any() and not d.any()
    import fpt_b200.css_cython as css_serial
    s, d = np.zeros(100), np.zeros(100)
    with pytest.raises(fpt.FptError) as e:
        css_serial.cluster_separation_scorer(av, bv, apos, bpos, 0, 50000, 2500, 500, av.size, bv.size, 10, 50, 0, 0, s, d)
    assert e.value.code == -3 and not s.any() and not d.any()


# ------------------------------------------------------------------------------------------------ whole pipeline
def test_pipeline_vcf_to_region_calls(fpt, oracle):
    """SURVEY 8(f) rows 1-4 around the CUDA path: VCF text -> native ingest -> Statistic stand-ins -> drop-in scorers (GPU)
    -> result files -> region callers, against the same pipeline with the oracle behind the drop-in argument lists."""
    from collections import OrderedDict
    import fpt_b200.synth as synth
    from fpt_b200 import ingest, results, tools
    asize, bsize, length, nsnp, seed = 6, 5, 70000, 1800, fpt.get_seed()
    names = ["a%d" % i for i in range(asize)] + ["b%d" % i for i in range(bsize)]
    gt = {3: "0/0", 0: "0|1", -3: "1/1", -128: "./."}
    lines = ["##fileformat=VCFv4.1", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names)]
    lens = OrderedDict()
    for c in range(2):
        ch = synth.chromosome(700 + c, length, nsnp, asize, bsize)
        lens["chr%d" % (c + 1)] = length
        codes = np.concatenate([ch["acodes"].reshape(-1, asize), ch["bcodes"].reshape(-1, bsize)], axis=1)
        for p, row in zip(ch["pos"], codes):
            lines.append("chr%d\t%d\t.\tA\tC\t.\t.\t.\tGT:DP\t%s" % (c + 1, p, "\t".join(gt[int(v)] + ":9" for v in row)))
    genome, info = ingest.read_vcf("\n".join(lines) + "\n", names[:asize], names[asize:])
    assert info["records"] == 2 * nsnp and list(genome) == list(lens)
    ta = OrderedDict((k, v.reference_layout()[0]) for k, v in genome.items())
    tb = OrderedDict((k, v.reference_layout()[1]) for k, v in genome.items())
    fet_text = tools.fisher_exact_test_snp_tool(ta, tb, lens, 2500, 500, 0.95, number=results.str_exact)
    css_text = tools.cluster_separation_score_tool(ta, tb, lens, mds=0, mc_treshold=10, mc_runs=200, number=results.str_exact)
    # the oracle behind the same host code
    want_fet, want_css = [results.FET_HEADER], [results.CSS_HEADER]
    for name in lens:
        a, b = ta[name], tb[name]
        n = length // 500
        s, d, cs, cp = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
        oracle.fpt_oracle_fet_scan(dptr(a.vals), dptr(b.vals), iptr(a.starts), iptr(b.starts), 0, length, 2500, 500, a.vals.size,
                                   b.vals.size, 0.95, dptr(s), dptr(d), 1, seed)
        oracle.fpt_oracle_css_scan(dptr(a.vals), dptr(b.vals), iptr(a.starts), iptr(b.starts), 0, length, 2500, 500, a.vals.size,
                                   b.vals.size, 10, 200, 0, 0, dptr(cs), dptr(cp), 1, seed)
        want_fet.append(results.format_windows(name, 500, s, d, results.str_exact))
        want_css.append(results.format_windows(name, 500, cs, cp, results.str_exact))
    for got, want, rtol in ((fet_text, "".join(want_fet), FET_RTOL), (css_text, "".join(want_css), CSS_RTOL)):
        gc, gs, g2, g3 = results.read_scan(got)
        wc, ws, w2, w3 = results.read_scan(want)
        assert gc == wc and np.array_equal(gs, ws) and len(gc) > 200        # same windows written, same order
        np.testing.assert_allclose(g2, w2, rtol=rtol, atol=1e-12)
        np.testing.assert_allclose(g3, w3, rtol=1e-9 if rtol == FET_RTOL else 0, atol=1e-12 if rtol == FET_RTOL else 0)
    # region calls from the GPU-written files equal those from the oracle-written files
    assert (results.significant_css_regions_file(css_text, 2500, lens, fdr=0.2)
            == results.significant_css_regions_file("".join(want_css), 2500, lens, fdr=0.2))
    assert (results.significant_css_regions_file(css_text, 2500, lens, num_top=10)
            == results.significant_css_regions_file("".join(want_css), 2500, lens, num_top=10))
    assert (results.filter_fisher_scores_file(fet_text, 2500, lens, 0.95, 75.0)
            == results.filter_fisher_scores_file("".join(want_fet), 2500, lens, 0.95, 75.0))


# ------------------------------------------------------------------------------------------------ chunked upload
@pytest.mark.parametrize("semantics", [0, 1])
def test_scans_with_chunked_upload_match_single_pass(fpt, oracle, monkeypatch, semantics):
    """the host entry points cut the genotype upload into chunks and score windows as their SNPs arrive; forced here on
    a small input (8 chunks of 4096 SNPs) and compared with the single-pass result and the oracle, both layouts"""
    asize, bsize, regend, wsize, wstep, nsnp = 9, 8, 900000, 2500, 500, 30000
    ch, (av, bv, apos, bpos) = _synth(321, regend, nsnp, asize, bsize)
    n = regend // wstep
    res = {}
    for tag, chunk in (("one", None), ("many", "4096")):
        if chunk is None:
            monkeypatch.delenv("FPT_UPLOAD_CHUNK_BYTES", raising=False)
        else:
            monkeypatch.setenv("FPT_UPLOAD_CHUNK_BYTES", chunk)
        f64 = fpt.fet_scan(av, bv, ch["pos"], asize, bsize, regend, wsize, wstep, 0.95, semantics=semantics, seed=77)
        i8 = fpt.fet_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 0.95, semantics=semantics, seed=77)
        css = fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 5, 60, mds=0, semantics=semantics, seed=77)
        res[tag] = (f64, i8, css)
    for a, b in zip(res["one"], res["many"]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    s_o, d_o = np.zeros(n), np.zeros(n)
    oracle.fpt_oracle_fet_scan(dptr(av), dptr(bv), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, av.size, bv.size, 0.95, dptr(s_o), dptr(d_o),
                               semantics, 77)
    s_g, d_g, wr = res["many"][0]
    assert np.array_equal(wr == 1, s_o != 0) and (wr == 1).sum() > 1000
    np.testing.assert_allclose(s_g, s_o, rtol=FET_RTOL, atol=1e-13)
    np.testing.assert_allclose(d_g, d_o, rtol=1e-9, atol=1e-12)


def test_page_locked_and_pageable_inputs_agree(fpt):
    """caller arrays in page-locked memory are copied from directly; pageable ones (plain numpy) go through the library's staging
    buffers, chunk by chunk on helper threads: same results bit for bit, for both scans and both layouts"""
    import torch
    asize, bsize, regend, wsize, wstep, nsnp = 9, 8, 900000, 2500, 500, 30000
    ch, (av, bv, apos, bpos) = _synth(322, regend, nsnp, asize, bsize)
    keep = []

    def pinned(x):
        t = torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
        keep.append(t)
        return t.numpy()
    big_a, big_b = np.tile(ch["acodes"], 1), np.tile(ch["bcodes"], 1)
    for a, b in ((av, bv), (big_a, big_b)):
        f0 = fpt.fet_scan(a, b, ch["pos"], asize, bsize, regend, wsize, wstep, 0.95, seed=5)
        f1 = fpt.fet_scan(pinned(a), pinned(b), pinned(ch["pos"]), asize, bsize, regend, wsize, wstep, 0.95, seed=5)
        c0 = fpt.css_scan(a, b, ch["pos"], asize, bsize, regend, wsize, wstep, 5, 40, mds=0, seed=5)
        c1 = fpt.css_scan(pinned(a), pinned(b), pinned(ch["pos"]), asize, bsize, regend, wsize, wstep, 5, 40, mds=0, seed=5)
        for x, y in zip(f0 + c0, f1 + c1):
            assert np.array_equal(x, y, equal_nan=True)
    assert (f0[2] == 1).sum() > 1000


def test_concurrent_host_calls_serialise(fpt):
    """two Python threads (ctypes drops the GIL) scanning at once: the host entry points share per-device staging buffers and
    must therefore take turns; results equal the ones of back-to-back calls"""
    import threading
    asize, bsize, regend, wsize, wstep, nsnp = 8, 8, 400000, 2500, 500, 12000
    ch, _ = _synth(55, regend, nsnp, asize, bsize)
    ch2, _ = _synth(56, regend, nsnp, asize, bsize)
    want = [fpt.fet_scan(c["acodes"], c["bcodes"], c["pos"], asize, bsize, regend, wsize, wstep, 0.95, seed=3) for c in (ch, ch2)]
    want_css = fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 10, 100, seed=3)
    got = {}

    def run(tag, fn):
        got[tag] = [fn() for _ in range(6)]

    th = [threading.Thread(target=run, args=("a", lambda: fpt.fet_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 0.95, seed=3))),
          threading.Thread(target=run, args=("b", lambda: fpt.fet_scan(ch2["acodes"], ch2["bcodes"], ch2["pos"], asize, bsize, regend, wsize, wstep, 0.95, seed=3))),
          threading.Thread(target=run, args=("c", lambda: fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 10, 100, seed=3)))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for tag, w in (("a", want[0]), ("b", want[1]), ("c", want_css)):
        assert len(got[tag]) == 6
        for res in got[tag]:
            for x, y in zip(res, w):
                assert np.array_equal(x, y)


@pytest.mark.parametrize("asize,bsize,nsnp", [(500, 500, 600), (200, 200, 100), (152, 148, 100)])
def test_css_large_cohort_mds_forms_agree(fpt, asize, bsize, nsnp):
    """Large-cohort classical MDS: the Lanczos product from 8-bit count codes squared arithmetically with the blanks as a list
    (default), from 8-bit count codes through the table of squares (m % 8 == 0), from 16-bit count codes (m % 4 == 0)
    and from the fp64 matrix B give the same embedding up to rounding — dense windows, and sparse ones where many pairs never
    differ and take the fill value — hence the same scores to 1e-9 and the same permutation p-values."""
    from fpt_b200 import api
    regend, wsize, wstep, seed = 200000, 50000, 50000, 21
    ch, _ = _synth(400 + nsnp, regend, nsnp, asize, bsize)
    out = []
    try:
        for form in (3, 2, 1, 0):
            api.set_lanczos_form(form)
            out.append(fpt.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], asize, bsize, regend, wsize, wstep, 20, 100, mds=0, seed=seed, probes=True))
    finally:
        api.set_lanczos_form(3)
    s3, p3, w3, pr3 = out[0]
    assert (w3 == 1).sum() >= 3
    for s_f, p_f, w_f, pr_f in out[1:]:
        assert np.array_equal(w_f, w3)
        np.testing.assert_allclose(s_f, s3, rtol=1e-9, atol=1e-12)
        assert np.array_equal(p_f, p3)
        np.testing.assert_allclose(pr_f["evals"], pr3["evals"], rtol=1e-10, atol=1e-9)
