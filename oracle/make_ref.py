#!/usr/bin/env python
"""Recipe for `oracle/_ref/`: the UNMODIFIED reference path, so that it can travel to the GPU box.

TEST INFRASTRUCTURE ONLY.  The reference is pure Python (there is nothing to compile), so "building" it means
taking the files of the hot path exactly as they lie under `/root/reference` and placing them in the
git-ignored directory `oracle/_ref/` (listed in `.gitignore`, NOT in `.gpurunignore`: like a built `.so`, it ships
with the snapshot but never enters the history).  `__graft_entry__.build()` calls this in the build container;
on the GPU box `/root/reference` is absent and the prebuilt `oracle/_ref/` is used as it is.

What is taken (byte for byte, SHA-256 of every file recorded in `oracle/_ref/MANIFEST.json`):

* `models/stereoanywhere/**.py`  - `StereoAnywhere.forward` (stereoanywhere.py:95-299), `CorrBlock1D`
  (corr.py:75-132), the update block (update.py), `utils/utils.py`, encoders / hourglass the forward needs;
* `mapreduce_v2/*.py`            - `TileWrapper` (tile_wrapper.py) and its presets.

Who may use it: `tests/` (the real-model EPE gate, tile parity), `bench.py --impl reference` and the
`cpu_baseline` leg (the reference's own `CorrBlock1D` timed on the host cores, kind "reference").  The product
(`stereoanywhere_b200/`) never imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("SA_REFERENCE_SOURCE", "/root/reference")
TREES = ["models/stereoanywhere", "mapreduce_v2"]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make(verbose: bool = False) -> str | None:
    """Refresh `oracle/_ref/` from the reference tree; returns its path, or None when neither exists."""
    if not os.path.isdir(os.path.join(SOURCE, "models", "stereoanywhere")):
        return DEST if os.path.exists(os.path.join(DEST, "MANIFEST.json")) else None
    manifest = {}
    for tree in TREES:
        for root, _dirs, files in os.walk(os.path.join(SOURCE, tree)):
            for name in sorted(files):
                if not name.endswith(".py"):
                    continue
                src = os.path.join(root, name)
                rel = os.path.relpath(src, SOURCE)
                dst = os.path.join(DEST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                digest = _sha(src)
                if not (os.path.exists(dst) and _sha(dst) == digest):
                    shutil.copyfile(src, dst)
                manifest[rel] = digest
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "kei312/stereoanywhere (unmodified files, sha256)", "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} files from {SOURCE}")
    return DEST


def verify() -> bool:
    """True iff every file of `oracle/_ref/` still has the digest recorded when it was taken from the reference."""
    mf = os.path.join(DEST, "MANIFEST.json")
    if not os.path.exists(mf):
        return False
    files = json.load(open(mf))["files"]
    return all(os.path.exists(os.path.join(DEST, rel)) and _sha(os.path.join(DEST, rel)) == d for rel, d in files.items())


if __name__ == "__main__":
    out = make(verbose=True)
    print(out if out else "no reference tree and no prebuilt oracle/_ref")
    sys.exit(0 if out else 1)
