"""Checkpoint formats either side of the path: inference and trainer
layouts, `model.` prefix repair, shape-inferred config, ONNX initializers, packed blob."""
import json

import numpy as np
import pytest
import torch

from conftest import load_weights
from mlff_distiller_b200 import checkpoint as ck
from mlff_distiller_b200.student_model import StudentForceField
from oracle import reference_loader

PARAMS = {"original": 427292, "tiny": 77203, "ultra_tiny": 21459}  # benchmarks/m6_summary_cuda.txt:15-17


def test_config_inferred_from_shapes(variant, weights):
    state, cfg = weights
    inferred = ck.infer_config(state, None)
    assert inferred.hidden_dim == cfg["hidden_dim"] and inferred.num_rbf == cfg["num_rbf"]
    assert inferred.num_interactions == cfg["num_interactions"] and inferred.max_z == 100
    assert inferred.cutoff == 5.0
    assert ck.num_parameters(inferred) == PARAMS[variant]


def test_blob_layout(variant, weights):
    state, cfg = weights
    c = ck.infer_config(state, cfg)
    blob = ck.pack_weights(state, c)
    assert blob.dtype == np.float32
    assert blob.size == sum(int(np.prod(s)) for s in ck.expected_keys(c).values())
    h = c.hidden_dim
    assert np.array_equal(blob[: 101 * h], state["embedding.weight"].reshape(-1))
    assert np.array_equal(blob[-1:], state["energy_head.4.bias"])
    bad = dict(state)
    bad["energy_head.0.weight"] = bad["energy_head.0.weight"][:, :-1]
    with pytest.raises(ValueError):
        ck.pack_weights(bad, c)


def test_state_dict_keys_match_reference_names(variant, weights):
    state, cfg = weights
    c = ck.infer_config(state, cfg)
    model = StudentForceField.from_state(state, c, "cpu")
    assert sorted(model.state_dict().keys()) == sorted(ck.expected_keys(c).keys())
    assert model.num_parameters() == PARAMS[variant]
    for k, v in state.items():
        assert np.array_equal(model.state_dict()[k].numpy(), v), k
    for attr in ("hidden_dim", "num_interactions", "num_rbf", "cutoff", "max_z", "use_torch_cluster"):
        assert hasattr(model, attr)


def test_save_load_round_trip_all_layouts(tmp_path, weights):
    state, cfg = weights
    c = ck.infer_config(state, cfg)
    model = StudentForceField.from_state(state, c, "cpu")
    p = tmp_path / "best_model.pt"
    model.save(p)
    raw = torch.load(p, weights_only=False)
    assert set(raw) == {"model_state_dict", "config", "num_parameters"}
    assert set(raw["config"]) == {"hidden_dim", "num_interactions", "num_rbf", "cutoff", "max_z", "use_torch_cluster"}
    again = StudentForceField.load(p)
    for k, v in model.state_dict().items():
        assert torch.equal(v, again.state_dict()[k])
    # trainer layout: `model.` prefix, config polluted with training keys, no model keys
    sd = {"model." + k: v for k, v in model.state_dict().items()}
    torch.save({"epoch": 3, "model_state_dict": sd, "config": {"learning_rate": 1e-3, "batch_size": 16},
                "optimizer_state_dict": {}}, tmp_path / "trainer.pt")
    st2, c2, meta = ck.load_any(tmp_path / "trainer.pt")
    assert c2.hidden_dim == c.hidden_dim and c2.num_interactions == c.num_interactions
    assert meta["epoch"] == 3 and all(not k.startswith("model.") for k in st2)
    # bare state_dict and npz
    torch.save(model.state_dict(), tmp_path / "bare.pt")
    assert ck.load_any(tmp_path / "bare.pt")[1].num_rbf == c.num_rbf
    with pytest.raises(FileNotFoundError):
        ck.load_any(tmp_path / "missing.pt")


@pytest.mark.skipif(not reference_loader.available(), reason="reference tree not present")
def test_onnx_reader_and_reference_round_trip(tmp_path):
    ref = reference_loader.REFERENCE_ROOT
    a, ca, _ = ck.load_any(ref / "models/original_model.onnx")
    b, cb, _ = ck.load_any(ref / "models/student_model.onnx")  # same weights, `model.` prefix
    assert ca == cb and set(a) == set(b)
    for k in a:
        assert np.array_equal(a[k], b[k])
    gold, _ = load_weights("original")
    for k in gold:
        assert np.array_equal(a[k], gold[k])
    # a checkpoint written by the REFERENCE class loads here, and vice versa
    mod = reference_loader.load_reference_module("student_model")
    ref_model = reference_loader.build_reference_model(a, ca)
    ref_model.save(tmp_path / "ref.pt")
    ours = StudentForceField.load(tmp_path / "ref.pt")
    for k, v in ref_model.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), k
    ours.save(tmp_path / "ours.pt")
    back = mod.StudentForceField.load(tmp_path / "ours.pt")
    for k, v in ours.state_dict().items():
        assert torch.equal(v, back.state_dict()[k]), k


def _export_like_reference(model, path):
    """A TorchScript archive laid out like scripts/export_to_torchscript.py:77-104 writes it: a
    traced two-argument wrapper whose ``model`` attribute holds the student's modules."""
    class SimpleWrapper(torch.nn.Module):
        def __init__(self, base_model):
            super().__init__()
            self.model = base_model

        def forward(self, atomic_numbers, positions):   # traceable stand-in graph (never executed by us)
            return self.model.embedding(atomic_numbers).sum() + positions.sum() * self.model.rbf.centers.sum()

    traced = torch.jit.trace(SimpleWrapper(model), (torch.tensor([1, 6, 8]), torch.zeros(3, 3)))
    torch.jit.save(traced, str(path))


def test_torchscript_archive_is_a_checkpoint_format(tmp_path, weights):
    """`use_jit=True, jit_path=...` points at a TorchScript archive (inference/ase_calculator.py:216-236):
    its parameters (``model.`` prefix) are read like any other checkpoint."""
    state, cfg = weights
    c = ck.infer_config(state, cfg)
    model = StudentForceField.from_state(state, c, "cpu")
    p = tmp_path / "student_jit.pt"
    _export_like_reference(model, p)
    assert ck.is_torchscript_archive(p)
    model.save(tmp_path / "plain.pt")
    assert not ck.is_torchscript_archive(tmp_path / "plain.pt")
    st, c2, meta = ck.load_any(p)
    assert meta["format"] == "torchscript" and c2 == c
    for k, v in model.state_dict().items():
        assert np.array_equal(st[k], v.numpy()), k
    again = StudentForceField.load(p)
    assert again.num_parameters() == model.num_parameters()


def test_energy_only_wrapper_signature_and_keys(weights):
    from mlff_distiller_b200.student_model import EnergyOnlyWrapper
    state, cfg = weights
    model = StudentForceField.from_state(state, ck.infer_config(state, cfg), "cpu")
    w = EnergyOnlyWrapper(model)
    assert set(w.state_dict()) == {"model." + k for k in model.state_dict()}
    with pytest.raises(RuntimeError, match="no CPU fallback"):   # evaluation is CUDA-only
        w(torch.tensor([8, 1, 1]), torch.zeros(3, 3))
