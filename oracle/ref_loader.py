"""Imports the reference's UNMODIFIED source (`/root/reference/mlift`) on top of the
`jax` / `mici` / `arviz` stand-ins of oracle/ref_shim.

TEST INFRASTRUCTURE ONLY.  `/root/reference` exists in the build container only (not on
the GPU box), so this is used (i) by tests/golden/make_ref_golden.py to write the
committed fixtures tests/golden/ref_*.npz and (ii) by tests/test_cpu_ref_pinning.py, which
skips when the tree is absent.  Nothing is copied: the modules are imported from where
they lie.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("MLIFT_REFERENCE_ROOT", "/root/reference")
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mlift"))


def load():
    """Returns the imported `mlift` package of the reference tree (cached)."""
    if not available():
        raise ImportError(f"{REFERENCE_ROOT}/mlift not present")
    import numpy as onp

    sys.dont_write_bytecode = True  # /root/reference is read-only

    # NumPy >= 2 dropped aliases the 2021-era reference still calls (mlift/prior.py:75,81)
    if not hasattr(onp, "product"):
        onp.product = onp.prod
    for p in (REFERENCE_ROOT, SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    jax = importlib.import_module("jax")
    assert "torch-shim" in jax.__version__, "a real jax shadows the stand-in"
    return importlib.import_module("mlift")


def example_model(name):
    """`mlift.example_models.<name>` of the reference tree."""
    load()
    return importlib.import_module(f"mlift.example_models.{name}")
