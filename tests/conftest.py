import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def reference_lib():
    """oracle/_ref/libsgtd_ref.so: the reference's own STDesc.cpp / cluster_manager.hpp compiled
    against the oracle/shim stand-in headers.  Built here when /root/reference exists, otherwise the
    prebuilt file that ships with the snapshot is used; skipped only if neither is there."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    ref.lib()
    return ref
