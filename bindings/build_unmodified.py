"""Builds the reference's four Cython modules UNMODIFIED — the .pyx files are taken byte for byte from the reference
tree (statistics/fisher/fisher_cython{,_parallel}.pyx, statistics/css/css_cython{,_parallel}.pyx) and compiled against
the reference's own headers — and links them with libfpt_fisher.so / libfpt_css.so, which export the reference's literal
`void threadcompute(...)` / `void compute(...)` (csrc/alias/*.c) and forward to libfpt_b200.so. This is the link line
of fisher_setup.sh / fisher_parallel_setup.sh / css_setup.sh / css_parallel_setup.sh with the object files and GSL
swapped for one -l flag; no source edit.

    python bindings/build_unmodified.py [reference statistics dir] [output dir]

Nothing from the reference is copied into the repository: the .pyx and Cython's generated C live in a temporary
directory, the built extension modules go to bindings/_unmodified/ (git-ignored; they travel to the GPU box like every
built .so).
"""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "fast-parallel-tools-for-genome-wide-analysis-of-genomic-divergence_b200")
MODULES = {"fisher_cython_parallel": ("fisher", "fpt_fisher"), "fisher_cython": ("fisher", "fpt_fisher"),
           "css_cython_parallel": ("css", "fpt_css"), "css_cython": ("css", "fpt_css")}


def build(ref_stats="/root/reference/statistics", out=os.path.join(HERE, "_unmodified")):
    import numpy
    from Cython.Build import cythonize
    from setuptools import Extension, setup
    os.makedirs(out, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="fpt_unmodified_")
    try:
        exts = []
        for name, (sub, lib) in MODULES.items():
            src = os.path.join(tmp, name + ".pyx")
            shutil.copyfile(os.path.join(ref_stats, sub, name + ".pyx"), src)     # Cython writes its .c next to the .pyx
            exts.append(Extension(name, [src], include_dirs=[numpy.get_include(), os.path.join(ref_stats, sub)],
                                  library_dirs=[LIBDIR], libraries=[lib], runtime_library_dirs=[LIBDIR],
                                  define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")]))
        setup(name="reference_cython_modules_on_libfpt", ext_modules=cythonize(exts, language_level=2, quiet=True),
              script_args=["build_ext", "--build-lib", out, "--build-temp", os.path.join(tmp, "obj")])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


if __name__ == "__main__":
    print(build(*sys.argv[1:3]))
