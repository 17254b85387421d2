"""Host logic of the per-model filter splines (csrc/spline_table.h), no GPU: the C++ builder that
mlffd_model_create runs is compiled stand-alone (tests/spline_host_check.cpp) and its spline is
compared with the directly evaluated FP64 filter of the reference
(src/mlff_distiller/models/student_model.py:249-255, 285-292, 318-322, 350) for every layer of the
three trained variants.  Stated bound: |f - spline| <= 2e-7, |f' - spline'| <= 1e-5 per Angstrom."""
import json
import shutil
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, VARIANTS, load_weights


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = tmp_path_factory.mktemp("spline") / "spline_host_check"
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "tests" / "spline_host_check.cpp")],
                   check=True)
    return exe


@pytest.mark.parametrize("variant_name", VARIANTS)
def test_spline_builder_interpolates_the_reference_filter(checker, variant_name, tmp_path):
    state, cfg = load_weights(variant_name)
    H, K, rc = cfg["hidden_dim"], cfg["num_rbf"], cfg["cutoff"]
    w = state["rbf.widths"].astype(np.float32)
    gam = (np.float32(1.0) / (w * w)).astype(np.float32)      # student_model.py:252, in FP32
    for l in range(cfg["num_interactions"]):
        p = f"interactions.{l}.message.rbf_to_scalar."
        blob = tmp_path / f"layer{l}.bin"
        with open(blob, "wb") as f:
            f.write(struct.pack("iif", H, K, rc))
            for a in (state["rbf.centers"], gam, state[p + "0.weight"], state[p + "0.bias"],
                      state[p + "2.weight"], state[p + "2.bias"]):
                f.write(np.ascontiguousarray(a, dtype=np.float32).tobytes())
        out = subprocess.run([str(checker), str(blob)], check=True, capture_output=True, text=True).stdout
        rep = json.loads(out)
        assert rep["max_val_err"] <= 2e-7, (l, rep)
        assert rep["max_der_err"] <= 1e-5, (l, rep)
        assert rep["amp"] > 0.5   # a real filter, not zeros


def test_quintic_basis_is_a_partition_of_unity_with_matching_derivative():
    """The closed-form recursion used on the host and in the kernels, restated in numpy."""
    def basis(u):
        prev = [np.ones_like(u)]
        for p in range(1, 6):
            cur = []
            for j in range(p + 1):
                left = prev[j - 1] if j >= 1 else 0.0
                right = prev[j] if j <= p - 1 else 0.0
                cur.append(((u + p - j) * left + (j + 1 - u) * right) / p)
            if p == 5:
                d = [(prev[j - 1] if j >= 1 else 0.0) - (prev[j] if j <= 4 else 0.0) for j in range(6)]
            prev = cur
        return np.stack(prev), np.stack(d)
    u = np.linspace(0.0, 1.0, 1001)
    b, db = basis(u)
    assert np.allclose(b.sum(0), 1.0, atol=1e-14) and np.allclose(db.sum(0), 0.0, atol=1e-13)
    assert np.allclose(b[:, 0], np.array([1, 26, 66, 26, 1, 0]) / 120.0)
    eps = 1e-6
    bp, _ = basis(u + eps)
    bm, _ = basis(u - eps)
    assert np.allclose((bp - bm) / (2 * eps), db, atol=1e-8)
