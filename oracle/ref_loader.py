"""ORACLE / TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference files on CPU, on top of the pure-torch shim in
oracle/shim.  Two roots:

  * /root/reference (authoring container only): tests/golden/make_golden.py pins the
    travelling oracle (oracle/igcn_oracle.py) against the reference's own outputs;
  * baseline/_ref (git-ignored, NOT gpurun-ignored: it travels to the GPU box): a verbatim
    copy of the path's files made by oracle/make_ref.py; `bench.py --impl reference`
    imports the reference's own train() from there (cpu_baseline.kind = "reference").
"""
import os
import sys
import types

REF_ROOT = os.environ.get("IGCN_REFERENCE_ROOT", "/root/reference")
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


TRAVEL_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def available(root=None) -> bool:
    return os.path.isfile(os.path.join(root or REF_ROOT, "kernel", "sgcn_img_snp.py"))


def load(root=None, with_train=False):
    """Returns a namespace with the reference's own classes/modules, imported from `root` (default /root/reference).
    One root per process: the modules are cached in sys.modules under the reference's own names."""
    root = root or REF_ROOT
    if not available(root):
        raise RuntimeError("reference tree not present at %s" % root)
    for p in (root, SHIM):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [SHIM, root]
    # bypass kernel/__init__.py (it imports datasets.py -> TU datasets -> real PyG)
    if "kernel" not in sys.modules or not getattr(sys.modules["kernel"], "_igcn_stub", False):
        pkg = types.ModuleType("kernel")
        pkg.__path__ = [os.path.join(root, "kernel")]
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
    if with_train:
        ns.train_eval = importlib.import_module("kernel.train_eval_sgcn_img_snps")   # train(): :511-548
    ns.root = root
    return ns
