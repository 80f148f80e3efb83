import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the C oracle is test infrastructure: (re)build it if missing or stale
    src = os.path.join(ROOT, "oracle", "crf_oracle.c")
    lib = os.path.join(ROOT, "oracle", "libcrforacle.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import OracleLib
    return OracleLib()


@pytest.fixture(scope="session")
def reflib():
    from oracle.binding import RefLib, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libcrfref.so not built (needs /root/reference; `make -C oracle ref`)")
    return RefLib()
