import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


AGENT_CASES = ["cheb_proj_matern_q3", "cheb_proj_gauss_q4", "hub_fid_q5", "hub_proj_ess_q3", "yzcx_proj_gauss_q4",
               "yzcx_fid_q2", "kyr_proj_matern_q4", "kyr_fid_q3"]
