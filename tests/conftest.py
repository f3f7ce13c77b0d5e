import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # the CPU oracle restatement is test infrastructure: build it on demand
    so = os.path.join(ROOT, "oracle", "libhabdec_oracle.so")
    if not os.path.exists(so):
        subprocess.run(["make", "libhabdec_oracle.so"], cwd=os.path.join(ROOT, "oracle"), check=False)
    lib = os.path.join(ROOT, "habdec_b200", "libhabdec_b200.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-j8"], cwd=os.path.join(ROOT, "habdec_b200", "csrc"), check=False)


@pytest.fixture(scope="session")
def oracle_kind():
    """'ref' (the compiled reference) when oracle/_ref exists, else the restatement."""
    from oracle import pyoracle as po
    return "ref" if po.available("ref") else "orc"
