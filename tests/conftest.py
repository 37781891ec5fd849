import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """the in-tree C-ABI library must exist for every test (CPU tests only load it / call host code)"""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLD, "meta.json")) as f:
        return json.load(f)


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


@pytest.fixture(scope="session")
def state_dict():
    from demfi_b200 import synth
    return synth.make_state_dict(0)


def case_inputs(cfg):
    from demfi_b200 import synth
    x = synth.make_frames(cfg["h"], cfg["w"], seed=0, batch=cfg["batch"], smooth=cfg["smooth"])
    t = torch.tensor(cfg["t"], dtype=torch.float32).reshape(cfg["batch"], 1)
    return x, t
