"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu`` and call the CUDA path through the C-ABI;
everything else runs on CPU (oracle vs golden vectors, host logic, ABI surface, gloo multi-process logic)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"
for p in (str(ROOT), str(ROOT / "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir() -> Path:
    return GOLDEN


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / f"ref_{name}.npz") as z:
        return {k: z[k] for k in z.files}


def phys_from(g: dict):
    import pyafv_b200 as afv
    r, A0, P0, KA, KP, Lam, delta = (float(v) for v in g["phys"])
    return afv.PhysicalParams(r=r, A0=A0, P0=P0, KA=KA, KP=KP, Lambda=Lam, delta=delta)


def oracle_kwargs(g: dict) -> dict:
    r, A0, P0, KA, KP, Lam, delta = (float(v) for v in g["phys"])
    L = float(g["L"]) if "L" in g else 0.0
    return dict(r=r, A0=g["A0"] if "A0" in g else A0, P0=P0, KA=KA, KP=KP, Lambda=Lam, delta=delta,
                periodic=L if L > 0 else None)


STATIC_CASES = ["init_d0", "init_d045", "varyA0_d0", "varyA0_d045", "seed407", "seed46", "debug200", "n1", "n2_touch",
                "n2_apart", "n3_collinear", "right_angle", "init_params2", "rand_dense", "rand_mid", "rand_sparse",
                "rand_squeezed", "pbc_dense", "pbc_sparse", "pbc_hex_varyA0"]


def rings_equal_up_to_rotation(off_a, typ_a, off_b, typ_b) -> bool:
    """edges_type per cell must agree up to a cyclic rotation (the ring start is implementation-defined)."""
    if not np.array_equal(off_a, off_b):
        return False
    for i in range(len(off_a) - 1):
        a = typ_a[off_a[i]:off_a[i + 1]].tolist()
        b = typ_b[off_b[i]:off_b[i + 1]].tolist()
        if a != b and not any(a[k:] + a[:k] == b for k in range(1, len(a))):
            return False
    return True
