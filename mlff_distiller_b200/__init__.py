"""B200-native energy+force path for the distilled PaiNN student (MLFF-Distiller hot path).

Public surface mirrors the reference (paths relative to /root/reference):
``StudentForceField`` (src/mlff_distiller/models/student_model.py:532) and
``StudentForceFieldCalculator`` (src/mlff_distiller/inference/ase_calculator.py:61).  Everything
below them is hand-written sm_100a CUDA behind the C-ABI declared in include/mlffd.h; there is
no CPU fallback -- construction fails loudly if the library or a GPU is missing.
"""
from .checkpoint import ModelConfig, load_any, pack_weights, read_onnx_initializers  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # Lazy: importing the package must not require the CUDA library (CPU build/ABI checks).
    if name in ("StudentForceField", "EnergyOnlyWrapper", "radius_graph"):
        from . import student_model
        return getattr(student_model, name)
    if name == "StudentForceFieldCalculator":
        from .ase_calculator import StudentForceFieldCalculator
        return StudentForceFieldCalculator
    raise AttributeError(name)
