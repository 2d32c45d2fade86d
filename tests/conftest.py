import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.bindings import Reference, have_reference
    if not have_reference():
        pytest.skip("compiled reference (oracle/_ref) not present")
    return Reference()


@pytest.fixture(scope="session")
def reference512():
    from oracle.bindings import Reference, have_reference
    if not have_reference() or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_ldpc_dec512.so")):
        pytest.skip("compiled AVX512 reference not present")
    try:
        import subprocess
        if "avx512bw" not in open("/proc/cpuinfo").read():
            pytest.skip("host has no AVX512")
    except OSError:
        pytest.skip("cannot read cpuinfo")
    return Reference(avx512=True)


@pytest.fixture(scope="session")
def ldpc():
    """The product library, initialised on cuda:0 (GPU tests only)."""
    from openairinterface5g_b200.ldpc import load_LDPClib
    return load_LDPClib()
