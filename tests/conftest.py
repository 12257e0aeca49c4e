import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def iic():
    import iic_b200
    return iic_b200


@pytest.fixture(scope="session")
def engine(iic):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return iic.Engine(iic.VIT_B_16, "cuda:0", operand_dtype="bf16")


@pytest.fixture(scope="session")
def engine_f16(iic):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return iic.Engine(iic.VIT_B_16, "cuda:0", operand_dtype="f16")
