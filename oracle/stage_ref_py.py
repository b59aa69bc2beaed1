"""Staging recipe for ``oracle/_ref_py``: the reference's own PYTHON modules of the render path, unmodified.

TEST INFRASTRUCTURE ONLY (same status as ``oracle/_ref``, the rebuilt reference CUDA extensions): the files are copied
byte for byte from ``/root/reference`` into the git-ignored ``oracle/_ref_py/`` so that they travel to the GPU box with
the repository snapshot (``/root/reference`` does not exist there) — they never enter the history and nothing on the
product path imports them.  ``tests/test_gpu_dropin.py`` runs them in a child process to prove the drop-in claim of
INTEGRATION.md: the reference's ``gridencoder/grid.py`` does ``import _gridencoder as _backend`` (grid.py:9-12) and finds
THIS repository's ``_gridencoder.py`` shim; likewise ``_shencoder`` / ``_freqencoder``.

    python -m oracle.stage_ref_py            # copies the files listed below, writes MANIFEST.json (sha256 per file)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref_py")
REFERENCE = os.environ.get("SANERF_REFERENCE", "/root/reference")

# the hot path's Python side: operator wrappers, encoder factory, activation, field, renderer
FILES = [
    "activation.py",
    "encoding.py",
    "gridencoder/__init__.py",
    "gridencoder/grid.py",
    "shencoder/__init__.py",
    "shencoder/sphere_harmonics.py",
    "freqencoder/__init__.py",
    "freqencoder/freq.py",
    "nerf/network.py",
    "nerf/renderer.py",
    "nerf/utils.py",          # the Trainer (train_step / train_one_epoch / save_checkpoint / load_checkpoint): tests/test_gpu_dropin.py
]


def stage(verbose: bool = True) -> dict:
    """Copy the files (when the reference checkout exists) and return {relative path: sha256}; on a box without the
    checkout, return the manifest of what was staged earlier (empty dict if nothing)."""
    manifest_path = os.path.join(OUT, "MANIFEST.json")
    if not os.path.isdir(REFERENCE):
        if os.path.exists(manifest_path):
            with open(manifest_path) as f:
                return json.load(f)
        return {}
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REFERENCE, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print(f"[oracle/_ref_py] staged {len(manifest)} reference modules", file=sys.stderr)
    return manifest


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT, rel)) for rel in FILES)


if __name__ == "__main__":
    m = stage()
    for k, v in sorted(m.items()):
        print(v[:16], k)
    sys.exit(0 if m else 1)
