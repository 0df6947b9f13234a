"""pytest configuration: the `gpu` marker and import paths.

`oracle/` (repo root) is the CPU checker; `ggml` (ggml-libtorch_b200/ggml) is the product package.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ggml-libtorch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from util import patch_gguf_iq3s_to_reference_table
    patch_gguf_iq3s_to_reference_table()          # IQ3_S: gguf-py follows the reference's codebook (see the helper)


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
