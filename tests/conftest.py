"""pytest configuration: the `gpu` marker, import paths and session-scoped checker libraries.

`-m "not gpu"`: the oracle against the reference's golden vectors and the committed fixtures,
host logic, the C-ABI export check, the hostsim logic check and the 2-rank gloo strip test.
`-m gpu`: the parity tests proper, all through libspb200.so's C ABI.
Nothing here reads /root/reference at run time (only `make ref`, when the mount exists, does).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ora  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The product library and the checkers are built once per session (no-ops when fresh)."""
    import __graft_entry__ as entry
    if not os.path.exists(entry.LIB) or not ora.have_port() or not os.path.exists(ora.HOSTSIM_LIB):
        entry.build()


@pytest.fixture(scope="session")
def port(_built):
    return ora.load_port()


@pytest.fixture(scope="session")
def port_dm(_built):
    return ora.load_port_dm()


@pytest.fixture(scope="session")
def ref(_built):
    if not ora.have_ref():
        pytest.skip("oracle/_ref not built (reference sources were never mounted here)")
    return ora.load_ref()


@pytest.fixture(scope="session")
def ref_dm(_built):
    if not ora.have_ref():
        pytest.skip("oracle/_ref not built (reference sources were never mounted here)")
    return ora.load_ref_dm()


@pytest.fixture(scope="session")
def hostsim(_built):
    return ora.load_hostsim()


@pytest.fixture(scope="session")
def sp(_built):
    from vk_cinematic_b200 import sp as _sp
    return _sp


@pytest.fixture(scope="session")
def gpu_sp(sp):
    """The ctypes mirror with a device selected; fails (never skips) when CUDA is missing, so a
    GPU run cannot go green on a fallback."""
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    assert sp.lib.sp_b200_Init(0) == 0
    return sp


def checkers(request):
    """Both CPU checkers when available (the port always is)."""
    out = [("port", ora.load_port())]
    if ora.have_ref():
        out.append(("ref", ora.load_ref()))
    return out
