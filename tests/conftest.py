import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, 'bear_b200', 'data')
YSD1 = os.path.join(DATA, 'ysd1_lag_5_file_0_preshuf.tsv')
SPARSE = os.path.join(DATA, 'ex_seqs_kmap_for_var_pred.csv')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    """Every test session starts from a built libbear_b200.so (nvcc cross-compiles without a GPU)."""
    from bear_b200 import build
    build.build()


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('test is marked gpu but no CUDA device is visible')
    torch.cuda.set_device(0)
    return torch.device('cuda', 0)
