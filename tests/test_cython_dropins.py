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
