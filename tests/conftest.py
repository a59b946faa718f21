import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    has_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))


def golden_manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_names(kind):
    return sorted(k for k, v in golden_manifest().items() if v["kind"] == kind)


def split_rows(rows, counts):
    out, o = [], 0
    for c in counts.tolist():
        out.append(torch.from_numpy(rows[o : o + c].copy()))
        o += c
    return out


@pytest.fixture()
def knobs():
    """Set thread-local test knobs of libcerb_post.so (include/cerb_post.h: cerb_debug_set); all are dropped afterwards."""
    from cerberusdet_b200 import _lib

    _lib.debug_reset()
    yield _lib.debug_set
    _lib.debug_reset()
