"""Checkpoint IO for the PaiNN student: torch checkpoints, ONNX initializers, packed blobs.

Mirrors the on-disk formats of the reference (citations relative to /root/reference):

* inference checkpoints written by ``StudentForceField.save``
  (src/mlff_distiller/models/student_model.py:1075-1099):
  ``{'model_state_dict', 'config': {hidden_dim, num_interactions, num_rbf, cutoff, max_z,
  use_torch_cluster}, 'num_parameters'}``;
* trainer checkpoints (src/mlff_distiller/training/trainer.py:457-476) whose ``config`` is a full
  training-config dump and whose keys may carry a ``model.`` prefix when the model was wrapped
  (scripts/fix_checkpoint.py:55-73 strips it);
* the ONNX exports (models/original_model.onnx, benchmarks/{tiny,ultra_tiny}_model.onnx) that carry
  every parameter as an FP32 ``raw_data`` initializer under its ``state_dict`` name.  The last
  layer's ``update.mixing_matrix`` is pruned there because it never reaches the energy
  (student_model.py:736 uses ``scalar_features`` only).

Missing hyper-parameters are inferred from tensor shapes exactly like
``StudentForceField.load`` does (student_model.py:1131-1162).

The C-ABI takes one flat little-endian FP32 blob; :func:`pack_weights` documents its order, which
is also declared in include/mlffd.h.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Mapping, Optional, Tuple, Union

import numpy as np

MODEL_PARAM_KEYS = (
    "hidden_dim", "num_interactions", "num_rbf", "cutoff", "max_z", "learnable_rbf",
    "use_torch_cluster",
)


@dataclass(frozen=True)
class ModelConfig:
    hidden_dim: int = 128
    num_interactions: int = 3
    num_rbf: int = 20
    cutoff: float = 5.0
    max_z: int = 118
    learnable_rbf: bool = False
    use_torch_cluster: bool = True

    def as_dict(self) -> Dict[str, object]:
        return {
            "hidden_dim": self.hidden_dim,
            "num_interactions": self.num_interactions,
            "num_rbf": self.num_rbf,
            "cutoff": self.cutoff,
            "max_z": self.max_z,
            "use_torch_cluster": self.use_torch_cluster,
        }


# --------------------------------------------------------------------------------------
# Minimal protobuf reader for ONNX initializers (no ``onnx`` package needed)
# --------------------------------------------------------------------------------------

def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) for one protobuf message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            val, pos = _varint(buf, pos)
        elif wire == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wire == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wire == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:  # groups are not used by ONNX
            raise ValueError(f"unsupported protobuf wire type {wire}")
        yield field, wire, val


def _parse_tensor(buf: bytes) -> Tuple[str, Optional[np.ndarray]]:
    dims, dtype, name, raw, floats, int64s = [], 0, "", None, [], []
    for field, wire, val in _fields(buf):
        if field == 1:  # dims (possibly packed)
            if wire == 0:
                dims.append(val)
            else:
                p = 0
                while p < len(val):
                    d, p = _varint(val, p)
                    dims.append(d)
        elif field == 2:
            dtype = val
        elif field == 8:
            name = val.decode("utf-8")
        elif field == 9:
            raw = val
        elif field == 4:  # float_data
            if wire == 2:
                floats.extend(struct.unpack(f"<{len(val) // 4}f", val))
            else:
                floats.append(struct.unpack("<f", val)[0])
        elif field == 7:  # int64_data
            if wire == 0:
                int64s.append(val)
            else:
                p = 0
                while p < len(val):
                    d, p = _varint(val, p)
                    int64s.append(d)
    np_dtype = {1: np.float32, 7: np.int64, 11: np.float64, 6: np.int32}.get(dtype)
    if np_dtype is None:
        return name, None
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(np_dtype).newbyteorder("<")).astype(np_dtype)
    elif floats:
        arr = np.asarray(floats, dtype=np_dtype)
    elif int64s:
        arr = np.asarray(int64s, dtype=np_dtype)
    else:
        arr = np.zeros(0, dtype=np_dtype)
    return name, arr.reshape(dims) if dims else arr.reshape(())


def read_onnx_initializers(path: Union[str, Path]) -> Dict[str, np.ndarray]:
    """All initializers of an ONNX file as numpy arrays keyed by name.

    ModelProto.graph is field 7, GraphProto.initializer field 5, TensorProto dims/data_type/
    name/raw_data fields 1/2/8/9.
    """
    data = Path(path).read_bytes()
    out: Dict[str, np.ndarray] = {}
    for field, wire, val in _fields(data):
        if field == 7 and wire == 2:
            for gfield, gwire, gval in _fields(val):
                if gfield == 5 and gwire == 2:
                    name, arr = _parse_tensor(gval)
                    if arr is not None:
                        out[name] = arr
    return out


# --------------------------------------------------------------------------------------
# state_dict handling
# --------------------------------------------------------------------------------------

def strip_prefix(state: Mapping[str, object], prefix: str = "model.") -> Dict[str, object]:
    """Drop a wrapper prefix from every key that has it (scripts/fix_checkpoint.py:55-73)."""
    if not any(k.startswith(prefix) for k in state):
        return dict(state)
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in state.items()}


def _to_numpy(v) -> np.ndarray:
    if isinstance(v, np.ndarray):
        return v
    if hasattr(v, "detach"):
        return v.detach().cpu().numpy()
    return np.asarray(v)


def state_dict_from_onnx(path: Union[str, Path]) -> Dict[str, np.ndarray]:
    """Model tensors from an ONNX export, filtered to state_dict names, prefix stripped."""
    init = strip_prefix(read_onnx_initializers(path))
    keep = ("embedding.", "rbf.", "interactions.", "energy_head.")
    return {k: np.ascontiguousarray(v) for k, v in init.items() if k.startswith(keep)}


def infer_config(state: Mapping[str, object], config: Optional[Mapping[str, object]] = None
                 ) -> ModelConfig:
    """Model hyper-parameters: explicit config keys win, the rest is inferred from shapes.

    Same rules as ``StudentForceField.load`` (student_model.py:1119-1162): hidden_dim and max_z
    from ``embedding.weight``, num_rbf from ``rbf.centers``, num_interactions from the set of
    ``interactions.<l>`` indices, cutoff defaults to 5.0.
    """
    cfg = {k: v for k, v in (config or {}).items() if k in MODEL_PARAM_KEYS}
    if "hidden_dim" not in cfg and "embedding.weight" in state:
        cfg["hidden_dim"] = int(_to_numpy(state["embedding.weight"]).shape[1])
    if "num_rbf" not in cfg and "rbf.centers" in state:
        cfg["num_rbf"] = int(_to_numpy(state["rbf.centers"]).shape[0])
    if "num_interactions" not in cfg:
        idx = {int(k.split(".")[1]) for k in state if k.startswith("interactions.")}
        if idx:
            cfg["num_interactions"] = len(idx)
    if "cutoff" not in cfg:
        cfg["cutoff"] = 5.0
    if "max_z" not in cfg and "embedding.weight" in state:
        cfg["max_z"] = int(_to_numpy(state["embedding.weight"]).shape[0]) - 1
    if "use_torch_cluster" not in cfg:
        cfg["use_torch_cluster"] = True
    cfg["cutoff"] = float(cfg["cutoff"])
    return ModelConfig(**cfg)


def is_torchscript_archive(path: Union[str, Path]) -> bool:
    """True for a ``torch.jit.save`` archive (zip with a ``constants.pkl`` record), the format
    ``use_jit=True, jit_path=...`` points at (inference/ase_calculator.py:216-236)."""
    import zipfile
    try:
        with zipfile.ZipFile(path) as z:
            return any(n.endswith("/constants.pkl") or n == "constants.pkl" for n in z.namelist())
    except (zipfile.BadZipFile, OSError):
        return False


def load_any(path: Union[str, Path]) -> Tuple[Dict[str, np.ndarray], ModelConfig, Dict[str, object]]:
    """Read a checkpoint in any supported format.

    Returns ``(state_dict as numpy, ModelConfig, raw metadata)``.  Supported: ``.onnx`` exports,
    ``.npz`` archives of state_dict tensors (optionally with ``__config__`` json), and torch
    pickles in the inference or trainer layout (``model_state_dict`` key, optional ``model.``
    prefix), a bare state_dict, or a TorchScript archive of the ``(Z, R) -> E`` wrapper.
    """
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(
            f"Checkpoint not found: {path}\n"
            f"Please ensure the model has been trained and checkpoint saved."
        )
    meta: Dict[str, object] = {}
    suffix = path.suffix.lower()
    if suffix == ".onnx":
        state = state_dict_from_onnx(path)
        config = None
    elif suffix == ".npz":
        import json
        with np.load(path, allow_pickle=False) as z:
            state = {k: np.ascontiguousarray(z[k]) for k in z.files if not k.startswith("__")}
            config = json.loads(str(z["__config__"])) if "__config__" in z.files else None
        state = strip_prefix(state)
    else:
        import torch
        if is_torchscript_archive(path):
            ckpt = torch.jit.load(str(path), map_location="cpu")
        else:
            try:   # tensors, containers and primitives only: nothing in the file gets to run code
                ckpt = torch.load(path, map_location="cpu", weights_only=True)
            except Exception:
                # like the reference (student_model.py:1116): trainer checkpoints may carry arbitrary
                # pickled objects (configs, optimizer / scheduler state); only for files the caller trusts
                ckpt = torch.load(path, map_location="cpu", weights_only=False)
        if isinstance(ckpt, torch.jit.ScriptModule):
            # TorchScript export of the (Z, R) -> E wrapper (scripts/export_to_torchscript.py:77-104):
            # the parameters live under the wrapper's ``model.`` attribute
            raw_state, config = ckpt.state_dict(), {}
            meta = {"format": "torchscript"}
            if not any(k.endswith("embedding.weight") for k in raw_state):
                raise ValueError(
                    f"{path}: TorchScript archive holds no parameters (frozen by "
                    "optimize_for_inference?); export it without freezing or pass the checkpoint")
        elif isinstance(ckpt, dict) and "model_state_dict" in ckpt:
            raw_state = ckpt["model_state_dict"]
            config = ckpt.get("config", {})
            meta = {k: v for k, v in ckpt.items() if k not in ("model_state_dict",)}
        elif isinstance(ckpt, dict) and "state_dict" in ckpt:
            raw_state, config = ckpt["state_dict"], ckpt.get("config", {})
        else:
            raw_state, config = ckpt, {}
        state = {k: _to_numpy(v) for k, v in strip_prefix(raw_state).items()}
        if not isinstance(config, dict):
            config = {}
    cfg = infer_config(state, config)
    return state, cfg, meta


def complete_state(state: Mapping[str, np.ndarray], cfg: ModelConfig) -> Dict[str, np.ndarray]:
    """Fill tensors an export may have pruned so the state_dict has all 35(+) reference keys.

    The only known pruned tensor is the last layer's ``update.mixing_matrix`` (dead for energy and
    forces); it is filled with zeros.  ``rbf.centers``/``rbf.widths`` are recreated with the
    reference formula (student_model.py:224-227) if absent.
    """
    out = {k: np.asarray(v) for k, v in state.items()}
    for l in range(cfg.num_interactions):
        key = f"interactions.{l}.update.mixing_matrix"
        if key not in out:
            out[key] = np.zeros((3, 3), dtype=np.float32)
    if "rbf.centers" not in out:
        out["rbf.centers"] = np.linspace(0.0, cfg.cutoff, cfg.num_rbf, dtype=np.float32)
    if "rbf.widths" not in out:
        out["rbf.widths"] = np.full(cfg.num_rbf, np.float32(cfg.cutoff / cfg.num_rbf), np.float32)
    return out


def expected_keys(cfg: ModelConfig):
    """state_dict keys and shapes of the reference module for ``cfg``."""
    H, K = cfg.hidden_dim, cfg.num_rbf
    keys = {
        "embedding.weight": (cfg.max_z + 1, H),
        "rbf.centers": (K,),
        "rbf.widths": (K,),
    }
    for l in range(cfg.num_interactions):
        p = f"interactions.{l}."
        keys[p + "message.rbf_to_scalar.0.weight"] = (H, K)
        keys[p + "message.rbf_to_scalar.0.bias"] = (H,)
        keys[p + "message.rbf_to_scalar.2.weight"] = (3 * H, H)
        keys[p + "message.rbf_to_scalar.2.bias"] = (3 * H,)
        keys[p + "update.update_mlp.0.weight"] = (H, 2 * H)
        keys[p + "update.update_mlp.0.bias"] = (H,)
        keys[p + "update.update_mlp.2.weight"] = (3 * H, H)
        keys[p + "update.update_mlp.2.bias"] = (3 * H,)
        keys[p + "update.mixing_matrix"] = (3, 3)
    keys["energy_head.0.weight"] = (H // 2, H)
    keys["energy_head.0.bias"] = (H // 2,)
    keys["energy_head.2.weight"] = (H // 4, H // 2)
    keys["energy_head.2.bias"] = (H // 4,)
    keys["energy_head.4.weight"] = (1, H // 4)
    keys["energy_head.4.bias"] = (1,)
    return keys


def pack_weights(state: Mapping[str, np.ndarray], cfg: ModelConfig) -> np.ndarray:
    """Flatten the state_dict into the blob ``mlffd_model_create`` expects (include/mlffd.h).

    Order: every tensor of :func:`expected_keys` in that order, row-major, FP32.  Linear weights
    keep torch's ``[out, in]`` layout; the library re-lays them out on the device.
    """
    full = complete_state(state, cfg)
    chunks = []
    for key, shape in expected_keys(cfg).items():
        if key not in full:
            raise KeyError(f"checkpoint is missing tensor '{key}'")
        arr = np.asarray(full[key], dtype=np.float32)
        if tuple(arr.shape) != tuple(shape):
            raise ValueError(f"tensor '{key}' has shape {arr.shape}, expected {shape}")
        chunks.append(np.ascontiguousarray(arr).reshape(-1))
    return np.concatenate(chunks).astype("<f4", copy=False)


def num_parameters(cfg: ModelConfig) -> int:
    """Trainable parameter count (buffers ``rbf.*`` excluded, as ``num_parameters()`` does)."""
    return sum(int(np.prod(s)) for k, s in expected_keys(cfg).items() if not k.startswith("rbf."))


def save_checkpoint(path: Union[str, Path], state: Mapping[str, np.ndarray], cfg: ModelConfig):
    """Write the reference's inference checkpoint layout (student_model.py:1085-1098)."""
    import torch
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    full = complete_state(state, cfg)
    sd = {k: torch.from_numpy(np.array(full[k], dtype=np.float32)) for k in expected_keys(cfg)}
    torch.save({"model_state_dict": sd, "config": cfg.as_dict(),
                "num_parameters": num_parameters(cfg)}, path)
