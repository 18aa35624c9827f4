"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference checkout for the GPU box.

    python -m oracle.stage_reference            (also called by __graft_entry__.build() when /root/reference exists)

The reference (labhamlet/wavjepa) is an application repository with a flat layout (`wavjepa/`, `hear_api/`,
`hear_configs/`, `data_modules/` next to `train.py`) and no build configuration:
`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref /root/reference` fails at metadata
generation ("Multiple top-level packages discovered in a flat-layout"), and most of its pinned dependencies
(pytorch-lightning, webdataset, hydra, torch 2.7.0 ...) are absent from this image anyway.  So the "install" is what
`pip install --target` would have produced for pure-Python packages: the package directories copied byte for byte into
`baseline/_ref/` (git-ignored: nothing of the reference enters this repository's history; NOT gpurun-ignored: the
copy travels to the GPU box, where /root/reference does not exist).  `oracle/ref_loader.py` then imports it there with
in-process stand-ins for the two missing imports, exactly as it does from /root/reference here.

A manifest with the sha256 of every staged file is written next to the copy so that "unmodified" can be checked.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(_REPO, "baseline", "_ref")
# the hot path's packages + what they import at module load
PACKAGES = ("wavjepa", "hear_api", "hear_configs", "data_modules")
FILES = ("utils.py", "LICENSE")


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def stage(src: str = "/root/reference", dest: str = DEST) -> str | None:
    """Copies the reference's hot-path packages into baseline/_ref/ (idempotent).  Returns dest, or None when the
    source checkout does not exist (the GPU box: the staged copy, if any, is used as is)."""
    if not os.path.isdir(os.path.join(src, "wavjepa")):
        return None
    manifest = {}
    os.makedirs(dest, exist_ok=True)
    for pkg in PACKAGES:
        s, d = os.path.join(src, pkg), os.path.join(dest, pkg)
        if not os.path.isdir(s):
            continue
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.wav", "*.ckpt", "*.pt"))
    for f in FILES:
        if os.path.exists(os.path.join(src, f)):
            shutil.copy2(os.path.join(src, f), os.path.join(dest, f))
    for root, _, files in os.walk(dest):
        for f in files:
            if f == "MANIFEST.json":
                continue
            p = os.path.join(root, f)
            rel = os.path.relpath(p, dest)
            manifest[rel] = _sha(p)
            ref_p = os.path.join(src, rel)
            assert os.path.exists(ref_p) and _sha(ref_p) == manifest[rel], f"staged copy of {rel} differs from the reference"
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": "labhamlet/wavjepa (unmodified copy, see oracle/stage_reference.py)", "files": manifest},
                  f, indent=0, sort_keys=True)
    return dest


if __name__ == "__main__":
    out = stage(*(sys.argv[1:2]))
    print(out if out else "reference checkout not found: nothing staged")
