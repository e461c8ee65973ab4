"""Shadow of the `torch_geometric` import name (2.0.2 surface the IG-GCN path touches): put `integration/shims` in front of
sys.path and the reference's UNMODIFIED model files (kernel/sgcn_img_snp.py:4-8, kernel/sgcn.py:4-8, batch.py:2-3) resolve their
PyG operators to the igcn_b200 kernels."""
from . import data, nn, utils  # noqa: F401

__version__ = "2.0.2"


def is_debug_enabled():          # batch.py:112
    return False
