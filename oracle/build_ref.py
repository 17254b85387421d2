#!/usr/bin/env python
"""ORACLE recipe (test infrastructure): pack the reference's own implementation of the path into
``oracle/_ref/reference_path.zip`` so that it travels to the GPU box.

Run from the repo root:  python oracle/build_ref.py        (``__graft_entry__.build()`` calls it too)

The reference path is two pure-Python files (torch + numpy only):
    src/mlff_distiller/models/student_model.py          forward, radius graph, autograd forces
    src/mlff_distiller/models/analytical_gradients.py   imported lazily by student_model.py:853, :975
They are read where they lie under ``/root/reference`` and written, unmodified, into a zip archive
together with two EMPTY package markers generated here (the reference's own ``__init__.py`` imports a
``data`` sub-package the snapshot lacks).  ``oracle/_ref/`` is git-ignored (a build output: nothing
of the reference enters the history) but not gpurun-ignored, so ``bench.py --impl reference`` and
the ``cpu_baseline`` leg time the reference's code itself on the GPU box's host cores
(``cpu_baseline.kind = "reference"``); without the archive they fall back to the restatement in
``oracle/painn_oracle.py`` (``kind = "port"``).  ``oracle/reference_loader.py`` imports from the
archive with zipimport.
"""
from __future__ import annotations

import hashlib
import json
import sys
import zipfile
from pathlib import Path

ORACLE = Path(__file__).resolve().parent
ARCHIVE = ORACLE / "_ref" / "reference_path.zip"
REFERENCE_ROOT = Path("/root/reference")
MEMBERS = ["src/mlff_distiller/models/student_model.py",
           "src/mlff_distiller/models/analytical_gradients.py"]
FIXED_DATE = (2020, 1, 1, 0, 0, 0)     # reproducible archive bytes


def manifest_of_tree() -> dict:
    return {m: hashlib.sha256((REFERENCE_ROOT / m).read_bytes()).hexdigest() for m in MEMBERS}


def manifest_of_archive() -> dict | None:
    if not ARCHIVE.exists():
        return None
    try:
        with zipfile.ZipFile(ARCHIVE) as z:
            return json.loads(z.read("MANIFEST.json"))["sha256"]
    except (KeyError, ValueError, zipfile.BadZipFile):
        return None


def build(force: bool = False) -> Path | None:
    """Write the archive when the reference tree is present; return its path (None when neither the
    tree nor an earlier archive exists -- the GPU box only ever uses the prebuilt file)."""
    if not all((REFERENCE_ROOT / m).exists() for m in MEMBERS):
        return ARCHIVE if ARCHIVE.exists() else None
    want = manifest_of_tree()
    if not force and manifest_of_archive() == want:
        return ARCHIVE
    ARCHIVE.parent.mkdir(parents=True, exist_ok=True)
    tmp = ARCHIVE.with_suffix(".tmp")
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        def put(name: str, data: bytes) -> None:
            z.writestr(zipfile.ZipInfo(name, FIXED_DATE), data)
        put("mlff_distiller/__init__.py", b"")
        put("mlff_distiller/models/__init__.py", b"")
        for m in MEMBERS:
            put(m[len("src/"):], (REFERENCE_ROOT / m).read_bytes())
        put("MANIFEST.json", json.dumps({"source": str(REFERENCE_ROOT), "sha256": want}, indent=1).encode())
    tmp.replace(ARCHIVE)
    return ARCHIVE


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out if out else "reference tree absent and no archive built earlier")
