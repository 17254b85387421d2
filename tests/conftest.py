import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
VARIANTS = ("original", "tiny", "ultra_tiny")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_weights(variant):
    with np.load(GOLDEN / f"weights_{variant}.npz") as z:
        state = {k: z[k] for k in z.files if not k.startswith("__")}
        cfg = json.loads(str(z["__config__"]))
    return state, cfg


def load_golden(variant):
    with np.load(GOLDEN / f"golden_{variant}.npz") as z:
        return {k: z[k] for k in z.files}


def golden_cases(gold):
    return sorted({k[: -len("_numbers")] for k in gold if k.endswith("_numbers")})


@pytest.fixture(scope="session", params=VARIANTS)
def variant(request):
    return request.param


@pytest.fixture(scope="session")
def weights(variant):
    return load_weights(variant)


@pytest.fixture(scope="session")
def golden(variant):
    return load_golden(variant)
