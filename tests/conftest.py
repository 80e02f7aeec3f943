import os
import sys
from pathlib import Path

# The seeded scenes (my_depthsplat_b200/scenes.py) and the camera glue go through MKL (matmul, einsum, inverse), whose
# code path -- and last bits -- depend on the host CPU.  The golden fixtures were generated on the portable path;
# asking for it before the first MKL call keeps bit-level fixtures comparable between hosts (test_golden.py falls
# back to tolerances when the inputs still differ).
os.environ.setdefault("MKL_CBWR", "COMPATIBLE")

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
