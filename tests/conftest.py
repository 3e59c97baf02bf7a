"""Shared fixtures.  Markers: `gpu` = needs a real B200 (run with -m gpu on the GPU box); everything else runs on CPU."""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"
REF_SRC = Path("/root/reference/src")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name: str):
        return np.load(GOLDEN / f"{name}.npz")
    return load


@pytest.fixture(scope="session")
def oracle():
    from oracle import codec_oracle
    return codec_oracle


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference package (only present in the build container)."""
    if not REF_SRC.exists():
        pytest.skip("reference sources not present (GPU box)")
    from oracle.gen_golden import import_reference
    return import_reference()


@pytest.fixture(scope="session")
def lib():
    from clip_neural_image_conpression_b200 import _lib
    return _lib.load()
