import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _have_gpu():
    try:
        from bioshell_b200 import _lib
        return _lib.lib().bsa_device_count() > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(ROOT, "tests", "golden", "ref_kats.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def oracle_matrices():
    """name -> (score441, aa_index256) parsed by the C oracle from the shipped NCBI text."""
    from bioshell_b200.scoring import SubstitutionMatrixList, ncbi_text
    from oracle import c_oracle
    return {n: c_oracle.parse_ncbi(ncbi_text(n)) for n in SubstitutionMatrixList.ALL}


@pytest.fixture(scope="session")
def ctx():
    from bioshell_b200 import Context
    c = Context(0)
    yield c
    c.close()
