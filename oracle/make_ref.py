#!/usr/bin/env python
"""Recipe: stage the UNMODIFIED reference package under `oracle/_ref/` so the GPU box can time it.

TEST / BENCH INFRASTRUCTURE (never imported by `vima_b200/`).  The reference (vimalabs/VIMA) is pure Python, so there
is nothing to compile: "building" it means putting its `vima/` package where `oracle/ref_shim.py` can import it on a
box that has no `/root/reference`.  `oracle/_ref/` is a build artefact exactly like a compiled `.so`:

    * git-ignored (reference sources never enter this repo's history),
    * NOT gpurun-ignored (it travels to the GPU box with the snapshot),
    * bytes-identical to `/root/reference/vima` -- `MANIFEST.sha256` lists every staged file with the digest of its
      source, and `verify()` re-checks the staged tree against that manifest before anything imports it.

    python oracle/make_ref.py            # stage (no-op when up to date)
    python oracle/make_ref.py --verify   # check the staged tree against its manifest

`bench.py --impl reference` and the `gpu_eager` leg run this staged package through the reference's own public API
(`VIMAPolicy.forward_obs_token / forward / forward_action_decoder / forward_action_token`); `kind` is "reference" when it
is present and "port" (the oracle restatement) otherwise.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = "/root/reference"
DST_ROOT = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(DST_ROOT, "MANIFEST.sha256")


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def _py_files(root: str):
    for d, _, files in sorted(os.walk(os.path.join(root, "vima"))):
        for f in sorted(files):
            if f.endswith(".py"):
                yield os.path.relpath(os.path.join(d, f), root)


def staged() -> bool:
    return os.path.isfile(MANIFEST) and os.path.isdir(os.path.join(DST_ROOT, "vima"))


def verify() -> bool:
    """True iff every file named in the manifest is present with the recorded digest (and nothing else is staged)."""
    if not staged():
        return False
    want = {}
    for line in open(MANIFEST):
        dig, rel = line.strip().split("  ", 1)
        want[rel] = dig
    have = set(_py_files(DST_ROOT))
    if have != set(want):
        return False
    return all(_sha(os.path.join(DST_ROOT, rel)) == dig for rel, dig in want.items())


def stage(force: bool = False) -> str:
    """Copies /root/reference/vima/**/*.py into oracle/_ref/vima (build container only). Returns the staged root."""
    if not os.path.isdir(os.path.join(SRC_ROOT, "vima")):
        if staged():
            return DST_ROOT
        raise RuntimeError(f"{SRC_ROOT}/vima is not present and nothing is staged under {DST_ROOT}")
    rels = list(_py_files(SRC_ROOT))
    digs = {rel: _sha(os.path.join(SRC_ROOT, rel)) for rel in rels}
    if not force and staged():
        cur = {}
        for line in open(MANIFEST):
            dig, rel = line.strip().split("  ", 1)
            cur[rel] = dig
        if cur == digs and verify():
            return DST_ROOT
    shutil.rmtree(os.path.join(DST_ROOT, "vima"), ignore_errors=True)
    for rel in rels:
        dst = os.path.join(DST_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC_ROOT, rel), dst)
    with open(MANIFEST, "w") as fh:
        for rel in rels:
            fh.write(f"{digs[rel]}  {rel}\n")
    assert verify()
    return DST_ROOT


if __name__ == "__main__":
    if "--verify" in sys.argv:
        ok = verify()
        print("staged reference verified" if ok else "staged reference missing or modified")
        sys.exit(0 if ok else 1)
    print(stage(force="--force" in sys.argv))
