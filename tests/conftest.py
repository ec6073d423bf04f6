import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with -m gpu)")
    config.addinivalue_line("markers", "slow: exhaustive checks (still CPU-only)")


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np

    gdir = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(gdir, "golden.json")) as f:
        meta = json.load(f)
    codec = np.load(os.path.join(gdir, "codec_cases.npz"))
    tr = np.load(os.path.join(gdir, "translate_cases.npz"))
    return {"meta": meta, "codec": codec, "translate": tr}
