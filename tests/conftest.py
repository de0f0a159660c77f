import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the installed copy of the unmodified reference (git-ignored) is what the compiler / mediator tests build their
    # object graphs with: on a fresh checkout next to the reference checkout, install it before collection
    try:
        import runpy
        runpy.run_path(os.path.join(ROOT, "baseline", "install_ref.py"))["install"]()
    except Exception as error:  # noqa: BLE001 - those tests then skip and say why
        sys.stderr.write(f"baseline/_ref not installed: {error}\n")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): oracle/libecmc_oracle.so behind oracle/oracle.py."""
    from oracle import oracle as orc
    orc.lib()
    return orc


@pytest.fixture(scope="session", autouse=True)
def native_library():
    """The built libraries are git-ignored: on a fresh checkout compile libecmc_b200.so (nvcc cross-compiles without a
    GPU) before the first test needs it. An existing library is left alone -- `__graft_entry__.build()` rebuilds stale ones."""
    from jellyfysh_b200 import build
    if not os.path.exists(build.LIBRARY):
        build.build(force=True)
