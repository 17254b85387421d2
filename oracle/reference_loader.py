"""ORACLE helper (container only): import the real reference model by file path.

``import mlff_distiller`` fails in the snapshot (``src/mlff_distiller/__init__.py:30`` imports a
``data`` sub-package that is missing, and ``ase`` is not installed), but
``models/student_model.py`` and ``models/analytical_gradients.py`` only need torch + numpy.  Empty
parent-package stubs make them importable under their real dotted names, which
``forward_with_analytical_forces`` needs (student_model.py:853, :975).

``/root/reference`` does not exist on the GPU box: this module is used by
``tests/golden/make_golden.py`` (run here, output committed) and by CPU tests that skip when the
reference tree is absent.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def available() -> bool:
    return (REFERENCE_ROOT / "src/mlff_distiller/models/student_model.py").exists()


def load_reference_module(name: str = "student_model"):
    """Return ``mlff_distiller.models.<name>`` from the reference tree."""
    if not available():
        raise FileNotFoundError("reference tree not present")
    src = REFERENCE_ROOT / "src" / "mlff_distiller"
    if "mlff_distiller" not in sys.modules or not hasattr(sys.modules["mlff_distiller"], "__path__"):
        pkg = types.ModuleType("mlff_distiller")
        pkg.__path__ = [str(src)]
        sys.modules["mlff_distiller"] = pkg
    if "mlff_distiller.models" not in sys.modules:
        sub = types.ModuleType("mlff_distiller.models")
        sub.__path__ = [str(src / "models")]
        sys.modules["mlff_distiller.models"] = sub
    return importlib.import_module(f"mlff_distiller.models.{name}")


def build_reference_model(state, cfg):
    """Instantiate the reference ``StudentForceField`` with the given numpy state_dict."""
    import torch
    mod = load_reference_module("student_model")
    model = mod.StudentForceField(hidden_dim=cfg.hidden_dim, num_interactions=cfg.num_interactions,
                                  num_rbf=cfg.num_rbf, cutoff=cfg.cutoff, max_z=cfg.max_z)
    sd = {k: torch.from_numpy(__import__("numpy").array(v)) for k, v in state.items()}
    model.load_state_dict(sd, strict=False)
    model.eval()
    return model
