"""Test configuration.

Markers: `gpu` -- needs a B200 (run with `-m gpu` on the GPU box); everything else runs on CPU.
The native libraries are built once per session (nvcc cross-compiles sm_100a without a GPU).
"""
from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
REFERENCE_XML = Path("/root/reference/Scenes.xml")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (B200)")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    from pibiti_b200 import build
    build.build_all()


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="session")
def golden_steps():
    return np.load(GOLDEN / "steps.npz")


@pytest.fixture(scope="session")
def golden_repo_scenes():
    return np.load(GOLDEN / "repo_scenes.npz")


@pytest.fixture(scope="session")
def golden_ref_scenes():
    return np.load(GOLDEN / "ref_scenes.npz")


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import oracle as orc
    return orc.load("port")


@pytest.fixture(scope="session")
def oracle_any():
    """The reference build when it is present (it travels to the GPU box prebuilt), else the port."""
    from oracle import oracle as orc
    return orc.load(None)


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
