"""torch_geometric.utils names the path imports (kernel/sgcn_img_snp.py:5)."""
from igcn_b200.pyg import to_dense_batch  # noqa: F401
