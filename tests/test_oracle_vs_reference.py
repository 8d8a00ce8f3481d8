"""The CPU oracle against the unmodified reference compiled into oracle/_ref (skipped where the reference tree
was never available, e.g. a box that only received the repo without oracle/_ref). Randomised inputs; every
comparison of arithmetic the oracle restates is bit-exact."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

import checkers
from checkers import RefMatrix, dptr, iptr

CODES = np.array([3.0, -3.0, 0.0, -10000.0])


def test_binomial_matches_reference(oracle, ref_fet):
    import math
    for n in range(0, 90):
        for k in range(0, n + 1):
            assert oracle.fpt_oracle_binomial(n, k) == ref_fet.binomial(n, k)
            if n <= 67:
                assert oracle.fpt_oracle_binomial(n, k) == math.comb(n, k)


def test_fet_all_tables_up_to_n30(oracle, ref_fet):
    """every 2x2 table with N <= 30: identical bits to the reference's fet()"""
    tmp = (C.c_int * 4)()
    n_checked = 0
    for n in range(31):
        for a in range(n + 1):
            for b in range(n + 1 - a):
                for c in range(n + 1 - a - b):
                    f, g = (C.c_int * 4)(a, b, c, n - a - b - c), (C.c_int * 4)(a, b, c, n - a - b - c)
                    assert oracle.fpt_oracle_fet_exact_domain(f)
                    assert oracle.fpt_oracle_fet_exact(f) == ref_fet.fet(g, tmp)
                    n_checked += 1
    assert n_checked == 46376


@settings(max_examples=300, deadline=None)
@given(st.integers(31, 67), st.data())
def test_fet_random_tables_up_to_n67(oracle, ref_fet, n, data):
    a = data.draw(st.integers(0, n))
    b = data.draw(st.integers(0, n - a))
    c = data.draw(st.integers(0, n - a - b))
    f, g, tmp = (C.c_int * 4)(a, b, c, n - a - b - c), (C.c_int * 4)(a, b, c, n - a - b - c), (C.c_int * 4)()
    if oracle.fpt_oracle_fet_exact_domain(f):
        assert oracle.fpt_oracle_fet_exact(f) == ref_fet.fet(g, tmp)


def test_log_mode_against_exact_rationals(oracle):
    """outside the reference's u64 domain: the log-space walk against exact rational arithmetic of the same
    two-tailed rule (first tail towards the minimum cell, strict `<` on the second tail, doubling on equal margins)"""
    from math import log10
    rng = np.random.default_rng(9)
    exact_rule = checkers.fet_exact_rule

    for _ in range(300):
        n1, n2 = int(rng.integers(40, 400)), int(rng.integers(40, 400))
        f = rng.uniform(0.05, 0.95)
        a, c = int(rng.binomial(n1, f)), int(rng.binomial(n2, min(0.99, max(0.01, f + rng.normal(0, 0.08)))))
        t = (a, n1 - a, c, n2 - c)
        got = oracle.fpt_oracle_fet_neglog10((C.c_int * 4)(*t))
        P = exact_rule(*t)
        want = 0.0 if P == 1 else -(log10(P.numerator) - log10(P.denominator))
        assert got == pytest.approx(want, rel=1e-9, abs=1e-10), t


def test_percentile_std_window_match_reference(oracle, ref_fet):
    rng = np.random.default_rng(2)
    for npos in (2, 3, 7, 25, 64, 200):
        x = rng.gamma(1.0, 1.0, npos)
        a, b = x.copy(), x.copy()
        for q in (0.95, 0.5, 0.0, 0.999):
            assert oracle.fpt_oracle_percentile(dptr(a), npos, q) == ref_fet.percentile(dptr(b), npos, q)
    for npos, state in ((11, 1), (37, 0xABCDEF012345), (100, 2 ** 48 - 1)):
        asize = bsize = 20
        av, bv = rng.choice(CODES, npos * asize, p=[.45, .3, .23, .02]), rng.choice(CODES, npos * bsize, p=[.3, .45, .23, .02])
        res = np.zeros(2)
        f, t = (C.c_int * 4)(), (C.c_int * 4)()
        samples, stds, fets = np.zeros(npos), np.zeros(100), np.zeros(npos)
        ref_fet.fisher_exact_test(dptr(res), dptr(av), dptr(bv), asize, bsize, npos, f, t, dptr(samples), dptr(stds), 100, dptr(fets),
                                  checkers.state_to_ushort3(state), 0.95)
        snp = np.zeros(npos)
        oracle.fpt_oracle_fet_per_snp(dptr(av), dptr(bv), npos, asize, bsize, None, dptr(snp))
        out = np.zeros(2)
        oracle.fpt_oracle_fet_window(dptr(snp), npos, 0.95, 100, state, dptr(out))
        assert np.array_equal(out, res)


@pytest.mark.parametrize("geom", [(2500, 500, 60000), (1000, 1000, 51000), (700, 300, 33333), (500, 500, 50000)])
def test_fet_scan_scores_match_reference_compute_and_threadcompute(oracle, ref_fet, geom):
    wsize, wstep, regend = geom
    rng = np.random.default_rng(wsize)
    nsnp, asize, bsize = 900, 12, 9
    pos = np.sort(rng.choice(regend, size=nsnp, replace=False)).astype(np.int32)
    av, bv = rng.choice(CODES, nsnp * asize, p=[.45, .3, .23, .02]), rng.choice(CODES, nsnp * bsize, p=[.3, .45, .23, .02])
    apos, bpos = np.repeat(pos, asize).astype(np.int32), np.repeat(pos, bsize).astype(np.int32)
    n = regend // wstep
    for threaded, fn in ((0, ref_fet.compute), (1, ref_fet.threadcompute)):
        s_r, d_r = np.zeros(n + 4), np.zeros(n + 4)
        with checkers.silence_stdout():
            fn(dptr(av), dptr(bv), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, av.size, bv.size, 0.95, dptr(s_r), dptr(d_r))
        s_o, d_o = np.zeros(n), np.zeros(n)
        assert oracle.fpt_oracle_fet_scan(dptr(av), dptr(bv), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, av.size, bv.size, 0.95,
                                          dptr(s_o), dptr(d_o), threaded, 3) == 0
        assert np.array_equal(s_o, s_r[:n])                    # same windows visited, same scores, bit for bit
        assert np.array_equal(d_o != 0, d_r[:n] != 0)


def test_css_functions_match_reference(oracle, ref_css):
    rng = np.random.default_rng(4)
    for asize, bsize, npos in ((20, 20, 60), (5, 9, 17), (2, 2, 9), (1, 3, 30)):
        m = asize + bsize
        f = np.clip(rng.beta(.5, .5, npos), .05, .95)
        enc = np.array([3., 0., -3.])
        av = enc[rng.binomial(2, f[:, None], size=(npos, asize))].ravel().copy()
        bv = enc[rng.binomial(2, (1 - f)[:, None], size=(npos, bsize))].ravel().copy()
        D = RefMatrix(m, m)
        ref_css.compare_all(dptr(av), dptr(bv), asize, bsize, npos, D.pp)
        D2 = np.zeros((m, m))
        oracle.fpt_oracle_compare_all(dptr(av), dptr(bv), asize, bsize, npos, dptr(D2))
        assert np.array_equal(D.a, D2)
        k1, k2 = ref_css.fill_averages(D.pp, m), oracle.fpt_oracle_fill_averages(dptr(D2), m)
        assert k1 == k2 and np.array_equal(D.a, D2)
        if not k1:
            continue
        X, B, Z, T, L, Q = RefMatrix(m, 2), RefMatrix(m, m), RefMatrix(m, m), RefMatrix(m, m), RefMatrix(2, 2), RefMatrix(m, 2)
        ref_css.cmds(D.pp, X.pp, 2, m, B.pp, Z.pp, T.pp, L.pp, Q.pp)
        X2, ev = np.zeros((m, 2)), np.zeros(3)
        oracle.fpt_oracle_cmds(dptr(D2), m, dptr(X2), dptr(ev))
        if np.isfinite(X.a).all() and np.isfinite(X2).all() and ev[1] - ev[2] > 1e-8 * ev[0]:
            d1, d2 = RefMatrix(m, m), np.zeros((m, m))
            ref_css.calc_dist(X.pp, d1.pp, m)
            oracle.fpt_oracle_calc_dist(dptr(X2), m, dptr(d2))
            assert np.allclose(d1.a, d2, rtol=1e-9, atol=1e-11)
        Xf = np.nan_to_num(X.a.copy())
        Xr = RefMatrix(m, 2, Xf)
        dist = RefMatrix(m, m)
        ref_css.calc_dist(Xr.pp, dist.pp, m)
        dist2 = np.zeros((m, m))
        oracle.fpt_oracle_calc_dist(dptr(Xf), m, dptr(dist2))
        assert np.array_equal(dist.a, dist2)
        at, bt = np.arange(asize, dtype=np.int32), np.arange(asize, m, dtype=np.int32)
        s1, s2 = ref_css.css(dist.pp, iptr(at), iptr(bt), asize, bsize), oracle.fpt_oracle_css(dptr(dist2), m, iptr(at), iptr(bt), asize, bsize)
        assert s1 == s2
        for tres, runs, state in ((10, 300, 77), (1000, 150, 0xFEEDFACE), (1, 50, 5)):
            t1, t2 = np.arange(m, dtype=np.int32), np.arange(m, dtype=np.int32)
            p1 = ref_css.significance_treshold(dist.pp, iptr(t1), asize, bsize, s1, tres, runs, checkers.state_to_ushort3(state))
            s64 = C.c_uint64(state)
            p2 = oracle.fpt_oracle_significance(dptr(dist2), m, iptr(t2), asize, bsize, s2, tres, runs, C.byref(s64), None, None)
            assert p1 == p2 and np.array_equal(t1, t2)
        Xs, Zs, Bs, Ds = RefMatrix(m, 2, Xf.copy()), RefMatrix(m, 2), RefMatrix(m, m), RefMatrix(m, m)
        sig1 = ref_css.smacof(D.pp, m, 2, Xs.pp, Zs.pp, Bs.pp, Ds.pp, 300, 1e-6)
        X3 = Xf.copy()
        sig2 = oracle.fpt_oracle_smacof(dptr(D2), m, dptr(X3), 300, 1e-6, None)
        assert sig1 == sig2 and np.array_equal(Xs.a, X3)


@pytest.mark.parametrize("mds", [0, 2])
def test_css_scan_scores_match_reference_compute(oracle, ref_css, mds):
    rng = np.random.default_rng(6 + mds)
    regend, wsize, wstep, nsnp, asize, bsize = 40000, 2500, 500, 800, 8, 8
    pos = np.sort(rng.choice(regend, size=nsnp, replace=False)).astype(np.int32)
    f = np.clip(rng.beta(.5, .5, nsnp), .05, .95)
    enc = np.array([3., 0., -3.])
    av = enc[rng.binomial(2, f[:, None], size=(nsnp, asize))].ravel().copy()
    bv = enc[rng.binomial(2, np.where((pos // 4000) % 2 == 1, 1 - f, f)[:, None], size=(nsnp, bsize))].ravel().copy()
    apos, bpos = np.repeat(pos, asize).astype(np.int32), np.repeat(pos, bsize).astype(np.int32)
    n = regend // wstep
    s_r, p_r = np.zeros(n + 4), np.zeros(n + 4)
    with checkers.silence_stdout():
        ref_css.compute(dptr(av), dptr(bv), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, av.size, bv.size, 3, 20, 0, mds, dptr(s_r), dptr(p_r))
    s_o, p_o = np.zeros(n), np.zeros(n)
    assert oracle.fpt_oracle_css_scan(dptr(av), dptr(bv), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, av.size, bv.size, 3, 20, 0, mds,
                                      dptr(s_o), dptr(p_o), 0, 1) == 0
    assert np.array_equal(p_o != 0, p_r[:n] != 0)
    ok = np.isfinite(s_r[:n])
    np.testing.assert_allclose(s_o[ok], s_r[:n][ok], rtol=1e-9, atol=1e-12)


def test_gsl_standin_eigensolver_matches_numpy(ref_css, oracle):
    """the GSL stand-in behind the compiled reference (oracle/gsl_shim) and the oracle's own Jacobi against
    numpy.linalg.eigh: classical-MDS distances of random configurations, including repeated and integer distances"""
    rng = np.random.default_rng(0)
    for m in (2, 3, 5, 17, 40, 41, 100):
        for rep in range(4):
            pts = rng.normal(size=(m, 3))
            D = np.sqrt(((pts[:, None] - pts[None]) ** 2).sum(-1))
            if rep == 3:
                D = np.round(D * 2)
            Dm = RefMatrix(m, m, D.copy())
            X, B, Z, T, L, Q = RefMatrix(m, 2), RefMatrix(m, m), RefMatrix(m, m), RefMatrix(m, m), RefMatrix(2, 2), RefMatrix(m, 2)
            ref_css.cmds(Dm.pp, X.pp, 2, m, B.pp, Z.pp, T.pp, L.pp, Q.pp)
            Xo, evo = np.zeros((m, 2)), np.zeros(3)
            oracle.fpt_oracle_cmds(dptr(np.ascontiguousarray(D)), m, dptr(Xo), dptr(evo))
            J = np.eye(m) - 1.0 / m
            ev, evec = np.linalg.eigh(-0.5 * J @ (D * D) @ J)
            order = np.argsort(ev)[::-1]
            if ev[order[1]] <= 0 or (m > 2 and ev[order[1]] - ev[order[2]] < 1e-8 * ev[order[0]]):
                continue
            Xn = evec[:, order[:2]] * np.sqrt(ev[order[:2]])
            dn = np.sqrt(((Xn[:, None] - Xn[None]) ** 2).sum(-1))
            for Xc in (X.a, Xo):
                dc = np.sqrt(((Xc[:, None] - Xc[None]) ** 2).sum(-1))
                assert np.abs(dc - dn).max() <= 1e-10 * dn.max()
