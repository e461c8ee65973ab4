"""Shadow of kernel/sgcn.py."""
from igcn_b200.sgcn_models import SGCN_GAT, SGCN_GCN  # noqa: F401
