import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: larger-than-a-few-seconds case")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def golden_cases():
    """EaBNet cases (tools/make_golden.py)"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith(("gag_", "resample_")))


def gag_golden_cases():
    """GaGNet post-filter cases (tools/make_golden_gag.py), without the wrapper case"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith("gag_") and "wrapper" not in f)


def load_golden(name):
    import numpy as np
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    rec = {k: z[k] for k in z.files}
    rec["cfg"] = eval(str(rec["cfg"]))          # dict literal written by tools/make_golden.py
    rec["B"], rec["L"], rec["variant"] = int(rec["B"]), int(rec["L"]), str(rec["variant"])
    return rec
