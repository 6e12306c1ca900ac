"""pytest configuration: registers the `gpu` marker, puts the repo root (oracle/) and the package
directory (clskd_b200) on sys.path and makes sure the C-ABI library is built (nvcc cross-compiles
for sm_100a without a GPU; the build is cached by a content stamp)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "speech-enhancement-clskd_b200")
for p in (ROOT, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    lib = os.path.join(PKG, "clskd_b200", "libclskd_sm100.so")
    if not os.path.exists(lib):
        sys.path.insert(0, PKG)
        import build as _build
        _build.build()
    return lib


@pytest.fixture
def emu(monkeypatch):
    """numpy model of the C ABI on host memory (tests/cabi_emu.py) for host-logic tests."""
    import cabi_emu
    return cabi_emu.install(monkeypatch)
