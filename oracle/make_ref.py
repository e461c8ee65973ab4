#!/usr/bin/env python
"""ORACLE / TEST INFRASTRUCTURE ONLY.  Makes baseline/_ref: a verbatim copy of the reference files on the hot path.

    python oracle/make_ref.py            # authoring container only (needs /root/reference)

The reference is a directory of Python scripts (no setup.py, nothing to compile), so "installing" it is copying the files
that `kernel/train_eval_sgcn_img_snps.py::train` (:511-548) imports: the import closure is taken from a real import of that
module on top of oracle/shim, every file is copied byte for byte to the same relative path under baseline/_ref/, and
MANIFEST.json lists the sha256 of each.  baseline/_ref is git-ignored (no reference source enters the history) and not
gpurun-ignored, so it travels to the GPU box with the snapshot; `bench.py --impl reference` imports the reference's own
train() from there and falls back to the oracle port when the directory is absent.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def make(verbose=True):
    from oracle import ref_loader
    src = ref_loader.REF_ROOT
    if not ref_loader.available(src):
        if verbose:
            print("make_ref: %s not present, nothing to do" % src)
        return None
    dst = ref_loader.TRAVEL_ROOT
    manifest_path = os.path.join(dst, "MANIFEST.json")
    ref_loader.load(src, with_train=True)
    real = os.path.realpath(src) + os.sep
    files = sorted({os.path.realpath(m.__file__) for m in list(sys.modules.values())
                    if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(real)})
    manifest = {}
    for f in files:
        rel = os.path.relpath(f, real)
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(f, out)
        manifest[rel] = hashlib.sha256(open(out, "rb").read()).hexdigest()
    json.dump(dict(source=src, files=manifest), open(manifest_path, "w"), indent=1)
    if verbose:
        print("make_ref: %d reference files -> %s" % (len(manifest), dst))
    return dst


if __name__ == "__main__":
    make()
