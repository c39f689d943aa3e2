import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    # CPU checkers (oracle restatement; oracle/_ref when /root/reference exists)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    so = os.path.join(ROOT, "xictools_b200", "libb200spice.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "xictools_b200", "csrc")])
