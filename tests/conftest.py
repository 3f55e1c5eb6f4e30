import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NMB = os.path.join(ROOT, "models", "nightmare_v3", "mjmodel.nmb")
REF_MODELS = "/root/reference/models"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def nmb_path():
    return NMB


@pytest.fixture(scope="session")
def compiled_model():
    from nightmare_rl_b200 import mjcf
    return mjcf.CompiledModel.load(NMB)


@pytest.fixture(scope="session")
def oracle_model():
    from oracle import oracle as O
    return O.OracleModel(NMB)
