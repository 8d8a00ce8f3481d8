"""Importable alias of the product package.

The package directory is named after the reference repository
(``fast-parallel-tools-for-genome-wide-analysis-of-genomic-divergence_b200``), which is not a valid
Python identifier; this stub makes ``import fpt_b200`` (and ``fpt_b200.fisher_cython_parallel`` etc.)
resolve to the modules living there. No code lives here.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "fast-parallel-tools-for-genome-wide-analysis-of-genomic-divergence_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
