"""pytest configuration: registers the ``gpu`` marker and shared fixtures."""
import os
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_addoption(parser):
    parser.addoption("--shuffle", type=int, default=None, metavar="SEED",
                     help="run the collected tests in a random order (order-dependence check: the library and the host layer keep no state between calls)")


def pytest_collection_modifyitems(config, items):
    seed = config.getoption("--shuffle")
    if seed is not None:
        import random
        random.Random(seed).shuffle(items)
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    inp = fx["inputs"]
    X, Y, Z = inp["tsdf_xyz"].shape
    # rebuild the reference's non-contiguous [1,1,Z,Y,X] view (get_tsdf.py:95-97)
    inp["tsdf_volume"] = inp["tsdf_xyz"].reshape(1, 1, X, Y, Z).permute(0, 1, 4, 3, 2)
    return fx


# render fixtures (map_*/track_*); mapper_rows.pt belongs to the Mapper-side rows (tests/test_mapper_oracle.py)
GOLDEN_CASES = sorted(f[:-3] for f in os.listdir(GOLDEN_DIR)
                      if f.endswith(".pt") and f.startswith(("map_", "track_"))) if os.path.isdir(GOLDEN_DIR) else []
