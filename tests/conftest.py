import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# The test-suite runs on seeded weights with no BPE vocabulary on disk: opt in (explicitly, as a deployment never would)
# to the two stand-ins clip_compat otherwise refuses - seeded random initialisation and the byte-level tokenizer.
os.environ.setdefault("IIC_ALLOW_RANDOM_INIT", "1")
os.environ.setdefault("IIC_ALLOW_STANDIN_TOKENIZER", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def iic():
    import iic_b200
    return iic_b200


@pytest.fixture(scope="session")
def engine(iic):
    """bf16 instantiation of the kernels for the operator-level tests (each compares a kernel with fp32 torch on the SAME
    16-bit operands, so the operand format only changes the output rounding step)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return iic.Engine(iic.VIT_B_16, "cuda:0", operand_dtype="bf16")


@pytest.fixture(scope="session")
def engine_default(iic):
    """the engine as a user gets it: the default operand dtype (fp16, _lib.DEFAULT_OPERAND_DTYPE)"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    eng = iic.Engine(iic.VIT_B_16, "cuda:0")
    assert eng.op_dtype == torch.float16
    return eng


@pytest.fixture(scope="session")
def engine_f16(iic):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return iic.Engine(iic.VIT_B_16, "cuda:0", operand_dtype="f16")
