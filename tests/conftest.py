import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SIGLIP_ARTEFACTS = os.path.join(GOLDEN, "siglip")  # copies of the reference's shipped head files (weights, not code)


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
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_scoring():
    return np.load(os.path.join(GOLDEN, "scoring_golden.npz"))


@pytest.fixture(scope="session")
def golden_heads():
    return np.load(os.path.join(GOLDEN, "heads_golden.npz"))


@pytest.fixture(scope="session")
def golden_backbone():
    return np.load(os.path.join(GOLDEN, "backbone_golden.npz"))


@pytest.fixture(scope="session")
def golden_freq_train():
    return np.load(os.path.join(GOLDEN, "freq_train_golden.npz"))


@pytest.fixture(scope="session")
def golden_decoder():
    return np.load(os.path.join(GOLDEN, "decoder_golden.npz"))


@pytest.fixture(scope="session")
def golden_preprocess():
    return np.load(os.path.join(GOLDEN, "preprocess_golden.npz"))


@pytest.fixture(scope="session")
def golden_gray():
    return np.load(os.path.join(GOLDEN, "gray_golden.npz"))


@pytest.fixture(scope="session")
def shipped():
    """The reference's shipped G1 artefacts (siglip/*: 7.4 KB of safetensors + 2 JSON + coral_bins.npy)."""
    import json

    from safetensors.torch import load_file

    return {
        "freq": load_file(os.path.join(SIGLIP_ARTEFACTS, "freq_mlp.safetensors")),
        "fusion": load_file(os.path.join(SIGLIP_ARTEFACTS, "fusion_head.safetensors")),
        "cuts": json.load(open(os.path.join(SIGLIP_ARTEFACTS, "coral_cutpoints.json"))),
        "temp": json.load(open(os.path.join(SIGLIP_ARTEFACTS, "coral_temp.json"))),
        "bins": np.load(os.path.join(SIGLIP_ARTEFACTS, "coral_bins.npy")),
        "dir": SIGLIP_ARTEFACTS,
    }
