import gzip
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the shared libraries (product + oracle) once per session if they are missing."""
    need = [os.path.join(ROOT, "repeatresolver_b200", "librr_maxcorr.so"),
            os.path.join(ROOT, "repeatresolver_b200", "librr_msagen.so"),
            os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as ge
        ge.build()
    yield


def golden_index():
    with open(os.path.join(GOLD, "index.json")) as f:
        return json.load(f)


def golden_msa(name):
    with gzip.open(os.path.join(GOLD, name + ".msa.gz"), "rb") as f:
        return f.read()


def golden_maxcorrs(name, cov):
    with gzip.open(os.path.join(GOLD, f"{name}.c{cov}.maxcorrs.gz"), "rb") as f:
        return f.read()


GOLDEN_CASES = [(n, int(c)) for n, covs in sorted(golden_index().items()) for c in sorted(covs)]
