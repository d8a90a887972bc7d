import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)



def have_ordering_library() -> bool:
    """COLAMD/AMD are SuiteSparse libraries the product loads at run time: its own build of them
    (slip_lu_b200/_deps, made by slip_lu_b200/build.py) or whatever SLIP_B200_ORDERING_LIB names."""
    from slip_lu_b200 import build as b
    return bool(os.environ.get("SLIP_B200_ORDERING_LIB")) or os.path.exists(b.ORDERING_SO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def product():
    import __graft_entry__ as entry
    entry.build()
    import slip_lu_b200
    return slip_lu_b200.lib()


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def reference(oracle):
    """The unmodified reference library (only where oracle/_ref was built)."""
    from slip_lu_b200 import capi
    if not os.path.exists(oracle.REF_SO):
        pytest.skip("oracle/_ref/libslip_ref.so not built (needs /root/reference)")
    return capi.SlipLib(oracle.REF_SO)


@pytest.fixture(scope="session")
def gpu(product):
    if product.dll.SLIP_B200_device_count() < 1:
        pytest.fail("no CUDA device visible: gpu tests must run on the B200 box")
    return product
