"""ORACLE / TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference files from /root/reference on CPU, on top of the
pure-torch shim in oracle/shim.  This only works in the
authoring container (the GPU box has no /root/reference); it is used by
tests/golden/make_golden.py to pin the travelling oracle (oracle/igcn_oracle.py)
against the reference's own outputs.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("IGCN_REFERENCE_ROOT", "/root/reference")
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "kernel", "sgcn_img_snp.py"))


def load():
    """Returns a namespace with the reference's own classes/modules."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    for p in (REF_ROOT, SHIM):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [SHIM, REF_ROOT]
    # bypass kernel/__init__.py (it imports datasets.py -> TU datasets -> real PyG)
    if "kernel" not in sys.modules or not getattr(sys.modules["kernel"], "_igcn_stub", False):
        pkg = types.ModuleType("kernel")
        pkg.__path__ = [os.path.join(REF_ROOT, "kernel")]
        pkg._igcn_stub = True
        sys.modules["kernel"] = pkg
    import importlib

    ns = types.SimpleNamespace()
    ns.sgcn_img_snp = importlib.import_module("kernel.sgcn_img_snp")
    ns.sgcn = importlib.import_module("kernel.sgcn")
    ns.go_model = importlib.import_module("kernel.go_model")
    ns.hp = importlib.import_module("sgcn_hyperparameters")
    ns.batch = importlib.import_module("batch")
    ns.dataloader = importlib.import_module("dataloader")
    ns.Data = importlib.import_module("torch_geometric.data").Data
    return ns
