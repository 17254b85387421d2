"""ctypes binding of libmlffd.so (C ABI in include/mlffd.h) and its in-tree nvcc build.

The library is built IN-TREE (mlff_distiller_b200/csrc/libmlffd.so) with
``nvcc -gencode arch=compute_100a,code=sm_100a`` so it travels with the repository snapshot to
the GPU box.  There is no CPU fallback anywhere: if the library is missing or no GPU is present
the product path raises.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from pathlib import Path
from typing import List, Optional

CSRC = Path(__file__).resolve().parent / "csrc"
# MLFFD_LIB: load another build of the same sources (kernel A/B experiments, tools/); the default is the in-tree library
LIB_PATH = Path(os.environ["MLFFD_LIB"]) if os.environ.get("MLFFD_LIB") else CSRC / "libmlffd.so"
SOURCES = ["mlffd.cu"]
HEADERS = ["common.cuh", "edge_features.cuh", "neighbor.cuh", "cell_list.cuh", "skin_list.cuh", "tile_gemm.cuh", "filter.cuh", "filter_umma.cuh", "umma_rows.cuh", "message.cuh", "message_pipe.cuh", "message_team.cuh", "message_spline.cuh", "spline_table.h",
           "update.cuh", "readout.cuh", "md.cuh", "../../include/mlffd.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

# every symbol include/mlffd.h declares
EXPORTS = [
    "mlffd_version", "mlffd_last_error", "mlffd_model_create", "mlffd_model_destroy",
    "mlffd_workspace_reserve", "mlffd_neighbor_list", "mlffd_export_edges",
    "mlffd_energy_forces", "mlffd_get_status", "mlffd_filter_table", "mlffd_debug_buffer",
    "mlffd_profile_enable", "mlffd_profile_read", "mlffd_stage_name",
    "mlffd_md_kick_drift", "mlffd_md_kick_energy", "mlffd_set_dense_fallback", "mlffd_virial",
    "mlffd_status_async", "mlffd_filter_spline", "mlffd_edge_features", "mlffd_rbf_cutoff", "mlffd_set_skin",
]
NUM_STAGES = 10

MLFFD_OK, MLFFD_EINVAL, MLFFD_ECUDA, MLFFD_ECAPACITY, MLFFD_ENOMEM = 0, -1, -2, -3, -4
PRECISIONS = {"fp32": 0, "tc": 1, "tc_bf16": 3, "tc_fp16": 4}
FILTER_MODES = {"spline": 0, "table": 1}
ABI_VERSION = 2


class MlffdConfig(ctypes.Structure):
    _fields_ = [("hidden_dim", ctypes.c_int32), ("num_rbf", ctypes.c_int32),
                ("num_interactions", ctypes.c_int32), ("max_z", ctypes.c_int32),
                ("cutoff", ctypes.c_float), ("precision", ctypes.c_int32),
                ("filter_mode", ctypes.c_int32)]


class MlffdStatus(ctypes.Structure):
    _fields_ = [("num_atoms", ctypes.c_int64), ("num_edges", ctypes.c_int64),
                ("num_pairs", ctypes.c_int64), ("edge_capacity", ctypes.c_int64),
                ("overflow", ctypes.c_int32), ("max_degree", ctypes.c_int32),
                ("overflow_events", ctypes.c_int64), ("tc_saturated", ctypes.c_int32),
                ("skin_rebuilds", ctypes.c_int32)]


class MlffdProfile(ctypes.Structure):
    _fields_ = [("launches", ctypes.c_int64), ("stage_launches", ctypes.c_int64 * NUM_STAGES),
                ("stage_ms", ctypes.c_double * NUM_STAGES)]


def nvcc_path() -> Optional[str]:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    return None


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [(CSRC / h).resolve() for h in HEADERS]
    return any(d.exists() and d.stat().st_mtime > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into csrc/libmlffd.so for sm_100a (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found; cannot build libmlffd.so")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB_PATH), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = subprocess.run(cmd, cwd=str(CSRC), capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({' '.join(cmd)}):\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return LIB_PATH


_lib = None


def load(build_if_missing: bool = False) -> ctypes.CDLL:
    """dlopen libmlffd.so and declare the prototypes of include/mlffd.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if build_if_missing:
            build()
        else:
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the energy+force path)")
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.mlffd_version.restype = ctypes.c_int
    lib.mlffd_version.argtypes = []
    lib.mlffd_last_error.restype = ctypes.c_char_p
    lib.mlffd_last_error.argtypes = [vp]
    lib.mlffd_model_create.restype = ctypes.c_int
    lib.mlffd_model_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int,
                                       ctypes.POINTER(MlffdConfig), vp, ctypes.c_size_t]
    lib.mlffd_model_destroy.restype = None
    lib.mlffd_model_destroy.argtypes = [vp]
    lib.mlffd_workspace_reserve.restype = ctypes.c_int
    lib.mlffd_workspace_reserve.argtypes = [vp, i64, i64, i64]
    lib.mlffd_neighbor_list.restype = ctypes.c_int
    lib.mlffd_neighbor_list.argtypes = [vp, vp, vp, i32, i64, vp, vp, vp]
    lib.mlffd_export_edges.restype = ctypes.c_int
    lib.mlffd_export_edges.argtypes = [vp, vp, i64, ctypes.POINTER(i64), vp]
    lib.mlffd_energy_forces.restype = ctypes.c_int
    lib.mlffd_energy_forces.argtypes = [vp, vp, vp, vp, i32, i64, vp, vp, vp, vp, vp]
    lib.mlffd_get_status.restype = ctypes.c_int
    lib.mlffd_get_status.argtypes = [vp, ctypes.POINTER(MlffdStatus)]
    lib.mlffd_status_async.restype = ctypes.c_int
    lib.mlffd_status_async.argtypes = [vp, vp, vp]
    lib.mlffd_filter_table.restype = ctypes.c_int
    lib.mlffd_filter_table.argtypes = [vp, i32, vp, i64, vp, vp, vp]
    lib.mlffd_filter_spline.restype = ctypes.c_int
    lib.mlffd_filter_spline.argtypes = [vp, i32, vp, i64, vp, vp, vp]
    lib.mlffd_edge_features.restype = ctypes.c_int
    lib.mlffd_edge_features.argtypes = [vp, vp, i64, ctypes.c_float, i32, vp, vp, vp, vp]
    lib.mlffd_rbf_cutoff.restype = ctypes.c_int
    lib.mlffd_rbf_cutoff.argtypes = [vp, i64, vp, i32, ctypes.c_float, ctypes.c_float, vp, vp]
    lib.mlffd_debug_buffer.restype = ctypes.c_int
    lib.mlffd_debug_buffer.argtypes = [vp, ctypes.c_char_p, i32, ctypes.POINTER(vp),
                                       ctypes.POINTER(i64), ctypes.POINTER(i32)]
    lib.mlffd_profile_enable.restype = ctypes.c_int
    lib.mlffd_profile_enable.argtypes = [vp, i32]
    lib.mlffd_profile_read.restype = ctypes.c_int
    lib.mlffd_profile_read.argtypes = [vp, ctypes.POINTER(MlffdProfile)]
    lib.mlffd_stage_name.restype = ctypes.c_char_p
    lib.mlffd_stage_name.argtypes = [i32]
    lib.mlffd_set_skin.restype = ctypes.c_int
    lib.mlffd_set_skin.argtypes = [vp, ctypes.c_float]
    lib.mlffd_set_dense_fallback.restype = ctypes.c_int
    lib.mlffd_set_dense_fallback.argtypes = [vp, i32]
    lib.mlffd_virial.restype = ctypes.c_int
    lib.mlffd_virial.argtypes = [vp, vp, i32, vp, vp]
    f64 = ctypes.c_double
    lib.mlffd_md_kick_drift.restype = ctypes.c_int
    lib.mlffd_md_kick_drift.argtypes = [vp, i64, vp, vp, vp, vp, f64, vp, vp]
    lib.mlffd_md_kick_energy.restype = ctypes.c_int
    lib.mlffd_md_kick_energy.argtypes = [vp, i64, vp, vp, vp, f64, vp, i32, vp, vp, i32, vp]
    if lib.mlffd_version() != ABI_VERSION:
        raise RuntimeError("libmlffd.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def exported_symbols() -> List[str]:
    """Dynamic symbols of the built library (via ctypes lookups, no compute calls)."""
    lib = ctypes.CDLL(str(LIB_PATH))
    return [s for s in EXPORTS if hasattr(lib, s)]


class MlffdError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libmlffd error {code}: {message}")
        self.code = code
