"""CPU-side checks of the boundary: libfpt_b200.so loads and exports every symbol include/fpt_b200.h declares, the
ctypes structs mirror the C structs, the drop-in modules validate arguments like the Cython originals, and without a
GPU every compute entry point fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fpt_b200.h")


@pytest.fixture(scope="module")
def lib():
    import fpt_b200._lib as L
    if not os.path.exists(L.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(os.path.dirname(L.LIB_PATH), "csrc")], check=True, capture_output=True)
    return L


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fpt_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_functions()
    assert len(names) >= 25
    handle = C.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), "libfpt_b200.so does not export %s" % n
        assert n in lib.SYMBOLS, "%s is declared in the header but not bound in _lib.py" % n
    for n in lib.SYMBOLS:
        assert n in names, "%s is bound in _lib.py but not declared in include/fpt_b200.h" % n


def test_struct_layouts_match_the_header(lib, tmp_path):
    """compile a probe against the real header and compare sizeof/offsetof with the ctypes mirrors"""
    probe = tmp_path / "probe.c"
    probe.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "fpt_b200.h"\nint main(void){'
                     'printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(fpt_genotypes), offsetof(fpt_genotypes, nsnp), offsetof(fpt_genotypes, bsize),'
                     'sizeof(fpt_scan_range), offsetof(fpt_scan_range, window_begin), offsetof(fpt_scan_range, states_init),'
                     'sizeof(fpt_css_probes), offsetof(fpt_css_probes, smacof_sigma));return 0;}')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(lib.Genotypes), lib.Genotypes.nsnp.offset, lib.Genotypes.bsize.offset, C.sizeof(lib.ScanRange),
            lib.ScanRange.window_begin.offset, lib.ScanRange.states_init.offset, C.sizeof(lib.CssProbes),
            lib.CssProbes.smacof_sigma.offset]
    assert got == want


def test_window_state_is_pure_host_arithmetic(lib, oracle):
    h = lib.load()
    for seed, w, s in ((0, 0, 0), (20261018, 42, 1), (2 ** 64 - 1, 10 ** 9, 0)):
        assert h.fpt_window_state(seed, w, s) == oracle.fpt_oracle_window_state(seed, w, s)
    h.fpt_set_seed(12345)
    assert h.fpt_get_seed() == 12345
    h.fpt_set_seed(20261018)


def test_no_device_means_error_not_fallback(lib):
    import fpt_b200.api as api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.FptError) as e:
        api.fet_tables(np.array([[1, 2, 3, 4]], dtype=np.int32))
    assert e.value.code == lib.FPT_ERR_NO_DEVICE
    import fpt_b200.css_cython_parallel as cp
    import fpt_b200.fisher_cython_parallel as fp
    av = np.zeros(8)
    pos = np.repeat(np.arange(4, dtype=np.int32), 2)
    with pytest.raises(api.FptError):
        fp.fisher_exact_tester(av, av, pos, pos, 0, 1000, 100, 100, 8, 8, 0.95, np.zeros(10), np.zeros(10))
    with pytest.raises(api.FptError):
        cp.cluster_separation_scorer(av, av, pos, pos, 0, 1000, 100, 100, 8, 8, 10, 100, 0, 0, np.zeros(10), np.zeros(10))


def test_product_never_references_the_oracle():
    """the shipped package must not import, link or open anything under oracle/ or tests/"""
    pkg = os.path.join(ROOT, "fast-parallel-tools-for-genome-wide-analysis-of-genomic-divergence_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")) or fn == "Makefile":
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "liboracle" not in txt and "oracle/" not in txt and "fpt_oracle" not in txt, fn
                assert "checkers" not in txt and "libref_" not in txt, fn
    out = subprocess.run(["ldd", os.path.join(pkg, "libfpt_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libref" not in out


def test_dropin_argument_checks_mirror_cython_buffer_typing():
    import fpt_b200.fisher_cython_parallel as fp
    ok64, ok32 = np.zeros(8), np.zeros(8, dtype=np.int32)
    out = np.zeros(10)
    with pytest.raises(ValueError):      # np.ndarray[np.float64_t, ndim=1] rejects float32
        fp.fisher_exact_tester(ok64.astype(np.float32), ok64, ok32, ok32, 0, 1000, 100, 100, 8, 8, 0.95, out, out)
    with pytest.raises(ValueError):      # int64 positions
        fp.fisher_exact_tester(ok64, ok64, ok32.astype(np.int64), ok32, 0, 1000, 100, 100, 8, 8, 0.95, out, out)
    with pytest.raises(ValueError):      # 2-d array
        fp.fisher_exact_tester(ok64.reshape(2, 4), ok64, ok32, ok32, 0, 1000, 100, 100, 8, 8, 0.95, out, out)
    with pytest.raises(ValueError):      # strided view: the reference would read garbage through .data
        fp.fisher_exact_tester(np.zeros(16)[::2], ok64, ok32, ok32, 0, 1000, 100, 100, 8, 8, 0.95, out, out)
    with pytest.raises(TypeError):
        fp.fisher_exact_tester(list(ok64), ok64, ok32, ok32, 0, 1000, 100, 100, 8, 8, 0.95, out, out)
    with pytest.raises(ValueError):      # outputs shorter than regend/wstep
        fp.fisher_exact_tester(ok64, ok64, ok32, ok32, 0, 1000, 100, 100, 8, 8, 0.95, np.zeros(3), out)


def test_extended_api_checks_caller_supplied_outputs():
    """api.fet_scan / api.css_scan hand `scores`, `stddev`, `p` to C by raw pointer: wrong dtype, stride, size or a read-only
    array must be refused before the call, not written through"""
    import fpt_b200.api as api
    a = np.zeros(8, dtype=np.int8)
    pos = np.arange(4, dtype=np.int32)
    for bad in (np.zeros(10, dtype=np.float32), np.zeros(20)[::2], np.zeros(3), np.zeros((2, 5))):
        with pytest.raises(ValueError):
            api.fet_scan(a, a, pos, 2, 2, 1000, 100, 100, 0.95, scores=bad)
        with pytest.raises(ValueError):
            api.css_scan(a, a, pos, 2, 2, 1000, 100, 100, 5, 10, p=bad)
    ro = np.zeros(10)
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        api.fet_scan(a, a, pos, 2, 2, 1000, 100, 100, 0.95, stddev=ro)


def test_synthetic_generators():
    import fpt_b200.synth as synth
    for gen in (synth.chromosome, synth.chromosome_fast):
        ch = gen(5, 200000, 3000, 7, 9)
        assert ch["pos"].dtype == np.int32 and np.all(np.diff(ch["pos"]) > 0) and ch["pos"].size == 3000
        assert set(np.unique(ch["acodes"])) <= {-128, -3, 0, 3} and ch["acodes"].size == 3000 * 7 and ch["bcodes"].size == 3000 * 9
        av, bv, apos, bpos = synth.reference_layout(ch)
        assert set(np.unique(av)) <= {-10000.0, -3.0, 0.0, 3.0} and apos.size == av.size and bpos.size == bv.size
        assert np.array_equal(apos[::7], ch["pos"])
    T = synth.coverage_tables(1, 1000)
    assert T.shape == (1000, 4) and T.min() >= 0 and (T[:, 0] + T[:, 1]).max() <= 500
