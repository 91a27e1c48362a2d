"""pytest configuration: the `gpu` marker and shared fixture loaders."""
import json
import os
import sys

import pytest

# engines of one process that wait for each other inside kernels (tests/test_sharded_gpu.py) need their streams on
# separate hardware queues; must be set before CUDA initialises
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    return load_golden
