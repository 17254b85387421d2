"""The C-ABI library builds, loads and exports every symbol include/mlffd.h declares; without a
GPU every entry point fails loudly (no CPU fallback).  No compute calls here."""
import ctypes
import re

import numpy as np
import pytest

from conftest import ROOT
from mlff_distiller_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.load()


def declared_symbols():
    text = (ROOT / "include" / "mlffd.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mlffd_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    declared = declared_symbols()
    assert declared == sorted(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_version_and_stage_names(lib):
    assert lib.mlffd_version() == _lib.ABI_VERSION == 2
    names = [lib.mlffd_stage_name(i).decode() for i in range(_lib.NUM_STAGES)]
    assert names[:4] == ["neighbor", "embedding", "filter", "message_fwd"]
    assert lib.mlffd_stage_name(99) == b""


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.MlffdConfig) == 28
    assert ctypes.sizeof(_lib.MlffdStatus) == 56
    assert ctypes.sizeof(_lib.MlffdProfile) == 8 + 8 * _lib.NUM_STAGES * 2


def test_bad_arguments_are_rejected_before_any_gpu_work(lib):
    h = ctypes.c_void_p()
    blob = np.zeros(16, dtype=np.float32)
    cfg = _lib.MlffdConfig(48, 20, 3, 100, 5.0, 0)  # unsupported hidden_dim
    rc = lib.mlffd_model_create(ctypes.byref(h), 0, ctypes.byref(cfg),
                                blob.ctypes.data_as(ctypes.c_void_p), blob.size)
    assert rc == _lib.MLFFD_EINVAL and b"hidden_dim" in lib.mlffd_last_error(None)
    cfg = _lib.MlffdConfig(128, 20, 3, 100, 5.0, 0)  # wrong blob length
    rc = lib.mlffd_model_create(ctypes.byref(h), 0, ctypes.byref(cfg),
                                blob.ctypes.data_as(ctypes.c_void_p), blob.size)
    assert rc == _lib.MLFFD_EINVAL and b"expected 427332" in lib.mlffd_last_error(None)
    cfg = _lib.MlffdConfig(128, 20, 3, 100, 5.0, 0, 7)  # unknown filter mode
    rc = lib.mlffd_model_create(ctypes.byref(h), 0, ctypes.byref(cfg),
                                blob.ctypes.data_as(ctypes.c_void_p), blob.size)
    assert rc == _lib.MLFFD_EINVAL and b"filter_mode" in lib.mlffd_last_error(None)
    assert lib.mlffd_workspace_reserve(None, 1, 1, 1) == _lib.MLFFD_EINVAL


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from conftest import load_weights
    from mlff_distiller_b200.checkpoint import infer_config
    from mlff_distiller_b200.engine import Engine
    from mlff_distiller_b200.student_model import StudentForceField
    state, cfg = load_weights("ultra_tiny")
    with pytest.raises(RuntimeError):
        Engine(state, infer_config(state, cfg), "cuda:0")
    model = StudentForceField.from_state(state, infer_config(state, cfg), "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.tensor([8, 1, 1]), torch.zeros(3, 3))
    # the library itself refuses as well
    lib = _lib.load()
    from mlff_distiller_b200.checkpoint import pack_weights
    c = infer_config(state, cfg)
    blob = pack_weights(state, c)
    h = ctypes.c_void_p()
    mc = _lib.MlffdConfig(c.hidden_dim, c.num_rbf, c.num_interactions, c.max_z, c.cutoff, 0)
    rc = lib.mlffd_model_create(ctypes.byref(h), 0, ctypes.byref(mc),
                                blob.ctypes.data_as(ctypes.c_void_p), blob.size)
    assert rc == _lib.MLFFD_ECUDA


def test_header_is_plain_c99():
    """The drop-in boundary is a C ABI: include/mlffd.h must compile as C99 (no C++ or torch types), warning-free,
    and a C translation unit that takes the address of every declared entry point must compile against it."""
    import shutil
    import subprocess
    import tempfile
    from pathlib import Path
    from mlff_distiller_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    header = Path(__file__).resolve().parents[1] / "include" / "mlffd.h"
    proc = subprocess.run([gcc, "-x", "c", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", str(header)],
                          capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    with tempfile.TemporaryDirectory() as tmp:
        src = Path(tmp) / "use_all.c"
        body = "\n".join(f"    p[{i}] = (const void*)(size_t)&{name};" for i, name in enumerate(_lib.EXPORTS))
        src.write_text(f'#include <stddef.h>\n#include "mlffd.h"\nconst void* p[{len(_lib.EXPORTS)}];\nvoid take(void) {{\n{body}\n}}\n')
        proc = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", str(header.parent), "-c", str(src), "-o", str(Path(tmp) / "use_all.o")],
                              capture_output=True, text=True)
        assert proc.returncode == 0, proc.stderr
