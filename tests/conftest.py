import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    import torch
    # ATen is only the yard-stick in these tests: make it exact fp32 (cuDNN convolutions default to TF32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def lib_built():
    """libtamgcn.so, built in-tree (nvcc cross-compiles for sm_100a without a GPU)."""
    from tam_gcn_b200 import build
    return build.build()
