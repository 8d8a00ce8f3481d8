"""The reference reaches its C through four small Cython modules; bindings/*.pyx are those modules with the extern swapped for
libfpt_b200.so. Builds them (Cython + gcc, no GPU needed) and checks that they import, keep the reference's function names and
positional signatures, and fail loudly without a device. On a GPU box the results must equal the ctypes modules'."""
import inspect
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {"fisher_cython_parallel": "fisher_exact_tester", "fisher_cython": "fisher_exact_tester",
         "css_cython_parallel": "cluster_separation_scorer", "css_cython": "cluster_separation_scorer"}


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    out = tmp_path_factory.mktemp("cython_dropins")
    r = subprocess.run([sys.executable, "setup.py", "build_ext", "--build-lib", str(out), "--build-temp", str(out / "tmp")],
                       cwd=os.path.join(ROOT, "bindings"), capture_output=True, text=True)
    for f in os.listdir(os.path.join(ROOT, "bindings")):
        if f.endswith(".c"):
            os.remove(os.path.join(ROOT, "bindings", f))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    sys.path.insert(0, str(out))
    mods = {n: __import__(n) for n in NAMES}
    sys.path.remove(str(out))
    return mods


def test_cython_modules_keep_the_reference_surface(built):
    fet_args = ["avals", "bvals", "apos", "bpos", "regstart", "regend", "wsize", "wstep", "alen", "blen", "perc", "scores", "stddev"]
    css_args = ["avals", "bvals", "apos", "bpos", "regstart", "regend", "wsize", "wstep", "alen", "blen", "treshold", "runs",
                "drosophila", "mds", "scores", "p"]
    for name, fn in NAMES.items():
        f = getattr(built[name], fn)
        want = fet_args if fn == "fisher_exact_tester" else css_args
        assert list(inspect.signature(f).parameters) == want
        import importlib
        py = importlib.import_module("fpt_b200." + name)                   # the ctypes twin has the same surface
        assert list(inspect.signature(getattr(py, fn)).parameters) == want


def test_cython_modules_fail_loudly_without_a_gpu(built):
    import fpt_b200.api as api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    a = np.zeros(8)
    p = np.repeat(np.arange(4, dtype=np.int32), 2)
    with pytest.raises(RuntimeError):
        built["fisher_cython_parallel"].fisher_exact_tester(a, a, p, p, 0, 1000, 100, 100, 8, 8, 0.95, np.zeros(10), np.zeros(10))
    with pytest.raises(ValueError):                                         # Cython buffer typing, as in the reference
        built["css_cython"].cluster_separation_scorer(a.astype(np.float32), a, p, p, 0, 1000, 100, 100, 8, 8, 5, 10, 0, 0, np.zeros(10), np.zeros(10))


@pytest.mark.gpu
def test_cython_modules_match_ctypes_modules(built):
    import fpt_b200.css_cython as cpy
    import fpt_b200.fisher_cython_parallel as fpy
    import fpt_b200.synth as synth
    ch = synth.chromosome(3, 120000, 3000, 20, 20)
    av, bv, apos, bpos = synth.reference_layout(ch)
    n = 120000 // 500
    s1, d1, s2, d2 = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    built["fisher_cython_parallel"].fisher_exact_tester(av, bv, apos, bpos, 0, 120000, 2500, 500, av.size, bv.size, 0.95, s1, d1)
    fpy.fisher_exact_tester(av, bv, apos, bpos, 0, 120000, 2500, 500, av.size, bv.size, 0.95, s2, d2)
    assert np.array_equal(s1, s2) and np.array_equal(d1, d2) and s1.any()
    c1, p1, c2, p2 = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    built["css_cython"].cluster_separation_scorer(av, bv, apos, bpos, 0, 120000, 2500, 500, av.size, bv.size, 10, 100, 0, 0, c1, p1)
    cpy.cluster_separation_scorer(av, bv, apos, bpos, 0, 120000, 2500, 500, av.size, bv.size, 10, 100, 0, 0, c2, p2)
    assert np.array_equal(c1, c2) and np.array_equal(p1, p2) and p1.any()


# ------------------------------------------------------------------------------------------------ unmodified reference modules
UNMODIFIED = os.path.join(ROOT, "bindings", "_unmodified")
PKG = os.path.join(ROOT, "fast-parallel-tools-for-genome-wide-analysis-of-genomic-divergence_b200")


def _unmodified_modules():
    import importlib
    sys.path.insert(0, UNMODIFIED)
    try:
        for n in NAMES:
            sys.modules.pop(n, None)
        return {n: importlib.import_module(n) for n in NAMES}
    finally:
        sys.path.remove(UNMODIFIED)
        for n in NAMES:
            sys.modules.pop(n, None)


def test_alias_libraries_export_the_reference_symbols():
    """libfpt_fisher.so / libfpt_css.so export exactly the names the reference's .pyx files bind
    (threadfisher.h:33, cFisher.h:11, threadcss.h:36, css.h:10): `threadcompute` and `compute`."""
    import ctypes as C
    for lib in ("libfpt_fisher.so", "libfpt_css.so"):
        h = C.CDLL(os.path.join(PKG, lib))
        for sym in ("threadcompute", "compute", "fpt_alias_status"):
            assert hasattr(h, sym), "%s lacks %s" % (lib, sym)


def test_unmodified_reference_pyx_links_against_the_alias_libraries():
    """the reference's four .pyx, byte for byte, cythonized against the reference's own headers and linked with the
    alias libraries instead of the reference's objects + GSL (bindings/build_unmodified.py)"""
    if not os.path.isdir("/root/reference/statistics"):
        pytest.skip("reference tree absent (the prebuilt modules are exercised by the GPU test)")
    sys.path.insert(0, os.path.join(ROOT, "bindings"))
    try:
        import build_unmodified
    finally:
        sys.path.remove(os.path.join(ROOT, "bindings"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bindings", "build_unmodified.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    mods = _unmodified_modules()
    for name, fn in NAMES.items():
        assert callable(getattr(mods[name], fn))
    import fpt_b200.api as api
    if api.device_count() == 0:                       # void entry points: the failure is recorded, the outputs stay zero
        import ctypes as C
        a = np.zeros(8)
        p = np.repeat(np.arange(4, dtype=np.int32), 2)
        s, d = np.zeros(10), np.zeros(10)
        mods["fisher_cython"].fisher_exact_tester(a, a, p, p, 0, 1000, 100, 100, 8, 8, 0.95, s, d)
        assert C.CDLL(os.path.join(PKG, "libfpt_fisher.so")).fpt_alias_status() == -5 and not s.any() and not d.any()


@pytest.mark.gpu
def test_unmodified_reference_modules_run_on_the_gpu():
    """the unmodified reference Cython modules (prebuilt by __graft_entry__.build(), shipped with the snapshot) produce the
    same numbers as the ctypes drop-ins: the reference's call surface really is intact"""
    if not os.path.isdir(UNMODIFIED) or len([f for f in os.listdir(UNMODIFIED) if f.endswith(".so")]) < 4:
        pytest.skip("bindings/_unmodified not built (needs the reference tree at build time)")
    import fpt_b200.css_cython as cser
    import fpt_b200.css_cython_parallel as cpar
    import fpt_b200.fisher_cython as fser
    import fpt_b200.fisher_cython_parallel as fpar
    import fpt_b200.synth as synth
    mods = _unmodified_modules()
    ch = synth.chromosome(3, 120000, 3000, 20, 20)
    av, bv, apos, bpos = synth.reference_layout(ch)
    n = 120000 // 500
    for name, twin in (("fisher_cython_parallel", fpar), ("fisher_cython", fser)):
        s1, d1, s2, d2 = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
        mods[name].fisher_exact_tester(av, bv, apos, bpos, 0, 120000, 2500, 500, av.size, bv.size, 0.95, s1, d1)
        twin.fisher_exact_tester(av, bv, apos, bpos, 0, 120000, 2500, 500, av.size, bv.size, 0.95, s2, d2)
        assert np.array_equal(s1, s2) and np.array_equal(d1, d2) and s1.any()
    for name, twin in (("css_cython_parallel", cpar), ("css_cython", cser)):
        c1, p1, c2, p2 = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
        mods[name].cluster_separation_scorer(av, bv, apos, bpos, 0, 120000, 2500, 500, av.size, bv.size, 10, 100, 0, 0, c1, p1)
        twin.cluster_separation_scorer(av, bv, apos, bpos, 0, 120000, 2500, 500, av.size, bv.size, 10, 100, 0, 0, c2, p2)
        assert np.array_equal(c1, c2) and np.array_equal(p1, p2) and p1.any()
