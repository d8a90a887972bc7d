"""slip_lu_b200 -- B200-native exact sparse LU (SLIP LU hot path) behind the SLIP_LU C interface.

The product is the C-ABI shared library ``slip_lu_b200/libslip_lu_b200.so`` (host C + CUDA
kernels for sm_100a, built by ``slip_lu_b200.build``).  This package only holds the build
recipe, the ctypes mirror of the interface and synthetic-input generators.  There is no
CPU fallback: :func:`lib` raises if the library has not been built.
"""
import os

from . import capi  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libslip_lu_b200.so")
_lib = None


def lib() -> "capi.SlipLib":
    """The product library bound through ctypes (loaded once)."""
    global _lib
    if _lib is None:
        _lib = capi.SlipLib(LIB_PATH)
    return _lib
