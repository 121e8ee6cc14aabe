import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    meta, arr = {}, {}
    for k in z.files:
        v = z[k]
        if k.startswith("meta_"):
            v = v.item() if v.ndim == 0 else v
            meta[k[5:]] = v
        else:
            arr[k] = v.item() if v.ndim == 0 else v
    return meta, arr


def golden_names(kind=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        name = os.path.basename(p)[:-4]
        if kind is None:
            out.append(name)
            continue
        z = np.load(p, allow_pickle=False)
        if str(z["meta_kind"]) in ([kind] if isinstance(kind, str) else kind):
            out.append(name)
    return out


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
