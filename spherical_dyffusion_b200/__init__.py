"""Importable alias of the ``spherical-dyffusion_b200/`` package directory.

The package directory carries the repository's (hyphenated) name, which Python cannot import
directly; this shim extends its own ``__path__`` to that directory so that
``import spherical_dyffusion_b200`` and ``from spherical_dyffusion_b200.sfnonet import ...`` work.
"""
import os as _os

_PKG_DIR = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "spherical-dyffusion_b200"))
__path__.append(_PKG_DIR)

from ._lib import LIB_PATH, SfnoLibraryError, lib, load_library  # noqa: E402,F401
from . import ops  # noqa: E402,F401  (torch.ops.sfno_b200.* custom-op layer)
from .harmonics import InverseRealSHT, RealSHT  # noqa: E402,F401
from .sfnonet import SpectralConvS2, SphericalFourierNeuralOperatorNet  # noqa: E402,F401

__all__ = ["SphericalFourierNeuralOperatorNet", "SpectralConvS2", "RealSHT", "InverseRealSHT", "ops", "lib", "load_library", "LIB_PATH",
           "SfnoLibraryError"]
