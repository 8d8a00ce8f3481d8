"""Builds the four Cython drop-ins against libfpt_b200.so — what replaces fisher_setup.sh / fisher_parallel_setup.sh /
css_setup.sh / css_parallel_setup.sh of the reference (icc + GSL + pthreads there; here only the one shared library).

    python bindings/setup.py build_ext --build-lib <dir>
"""
import os

import numpy
from Cython.Build import cythonize
from setuptools import Extension, setup

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "fast-parallel-tools-for-genome-wide-analysis-of-genomic-divergence_b200")
NAMES = ["fisher_cython_parallel", "fisher_cython", "css_cython_parallel", "css_cython"]

ext = [Extension(n, [os.path.join(HERE, n + ".pyx")], include_dirs=[numpy.get_include(), os.path.join(ROOT, "include")],
                 library_dirs=[LIBDIR], libraries=["fpt_b200"], runtime_library_dirs=[LIBDIR],
                 define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")]) for n in NAMES]
setup(name="fpt_b200_cython_dropins", ext_modules=cythonize(ext, language_level=2, quiet=True), script_args=None)
