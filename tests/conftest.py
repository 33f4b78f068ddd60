import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    import torch
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    from oracle.resepformer_oracle import OracleSepformerSeparation
    return OracleSepformerSeparation(seed=0)


@pytest.fixture(scope="session")
def sds(oracle):
    return oracle.component_state_dicts()


@pytest.fixture(scope="session")
def cuda_lib_built():
    from clearconverse_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from clearconverse_b200.build import build
        build()
    return _lib.load_library()
