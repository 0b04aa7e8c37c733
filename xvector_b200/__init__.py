"""Importable alias for the product package, whose directory is named after the upstream
repo (``x-vector-kaldi-tf_b200/`` -- not a valid Python identifier).  ``import xvector_b200``
and ``import xvector_b200.models`` resolve to the files in that directory."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "x-vector-kaldi-tf_b200")
__path__.insert(0, _PKG_DIR)
PACKAGE_DIR = _PKG_DIR
