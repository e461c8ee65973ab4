"""Shadow of the `torch_scatter` import name (kernel/go_model.py:20)."""
from igcn_b200.pyg import scatter, scatter_add  # noqa: F401
