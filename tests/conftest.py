import os
import sys

import pytest

# Ranks emulated on ONE GPU (tests/test_gpu_sharded.py) spin on each other from different streams: every
# stream needs its own hardware work queue, or a spinning kernel can sit in front of the kernel it waits
# for. The default is 8 queues; 8 emulated ranks plus the operator stream need more. (Read at context
# creation; irrelevant to the product path, where each rank owns its GPU.)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
