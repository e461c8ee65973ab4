"""torch_geometric.nn names the path imports (kernel/sgcn_img_snp.py:4,8; kernel/sgcn.py:4,8)."""
from igcn_b200.pyg import GATConv, GCNConv, global_add_pool, global_max_pool, global_mean_pool  # noqa: F401


def _not_on_path(name):
    def _raise(*a, **k):
        raise RuntimeError("torch_geometric.nn.%s is imported by the reference but never called on the IG-GCN hot path; "
                           "igcn_b200 does not implement it" % name)
    return _raise


ChebConv = _not_on_path("ChebConv")                    # imported at kernel/sgcn.py:4, used only by the commented-out variants
global_sort_pool = _not_on_path("global_sort_pool")
