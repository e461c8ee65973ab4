"""ORACLE ONLY. torch_geometric.utils 2.0.2 restated in plain torch."""
import torch


def add_remaining_self_loops(edge_index, edge_attr, fill_value, num_nodes):
    """Drop existing self loops from the edge list, append one loop per node LAST;
    a node that already had a loop keeps that loop's weight (last one wins),
    everyone else gets `fill_value` (PyG 2.0.2 utils/loop.py)."""
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    loops = torch.arange(num_nodes, dtype=src.dtype, device=src.device)
    if edge_attr is not None:
        loop_attr = edge_attr.new_full((num_nodes,) + tuple(edge_attr.shape[1:]), float(fill_value))
        is_loop = ~keep
        # index_put with duplicate indices: torch CPU writes in order, last wins
        loop_attr = loop_attr.index_put((src[is_loop],), edge_attr[is_loop])
        edge_attr = torch.cat([edge_attr[keep], loop_attr], 0)
    edge_index = torch.cat([edge_index[:, keep], loops.unsqueeze(0).repeat(2, 1)], 1)
    return edge_index, edge_attr


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    """Append one loop per node. fill_value may be a float or a reduce name
    ('mean','add',...) applied to the incoming edge attrs of each node."""
    n = num_nodes
    loops = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device).unsqueeze(0).repeat(2, 1)
    if edge_attr is not None:
        tail = tuple(edge_attr.shape[1:])
        if fill_value is None:
            loop_attr = edge_attr.new_full((n,) + tail, 1.0)
        elif isinstance(fill_value, (int, float)):
            loop_attr = edge_attr.new_full((n,) + tail, float(fill_value))
        elif isinstance(fill_value, str):
            dst = edge_index[1]
            acc = edge_attr.new_zeros((n,) + tail).index_add(0, dst, edge_attr)
            if fill_value == "mean":
                cnt = edge_attr.new_zeros(n).index_add(0, dst, edge_attr.new_ones(dst.numel()))
                cnt = cnt.clamp(min=1)
                acc = acc / (cnt.view((-1,) + (1,) * len(tail)))
            elif fill_value not in ("add", "sum"):
                raise ValueError(fill_value)
            loop_attr = acc
        else:
            raise ValueError(fill_value)
        edge_attr = torch.cat([edge_attr, loop_attr], 0)
    return torch.cat([edge_index, loops], 1), edge_attr


def segment_softmax(src, index, num_nodes):
    """PyG utils.softmax: per-target max-shifted softmax, eps 1e-16 in the denominator."""
    shape = (num_nodes,) + tuple(src.shape[1:])
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    mx = torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device)
    mx = mx.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    out = (src - mx.detach()[index]).exp()
    den = torch.zeros(shape, dtype=src.dtype, device=src.device).index_add(0, index, out)
    return out / (den[index] + 1e-16)


softmax = segment_softmax


def to_dense_batch(x, batch=None, fill_value=0.0, max_num_nodes=None, batch_size=None):
    """(N,F) -> (B, Nmax, F) + mask; nodes keep their in-graph order
    (reference call sites kernel/sgcn_img_snp.py:226,265,294)."""
    if batch is None:
        return x.unsqueeze(0), torch.ones(1, x.size(0), dtype=torch.bool, device=x.device)
    b = int(batch.max().item()) + 1 if batch_size is None else batch_size
    counts = torch.zeros(b, dtype=torch.long, device=x.device).index_add(0, batch, torch.ones_like(batch))
    start = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    nmax = int(counts.max().item()) if max_num_nodes is None else max_num_nodes
    pos = torch.arange(batch.numel(), device=x.device) - start[batch] + batch * nmax
    out = x.new_full((b * nmax,) + tuple(x.shape[1:]), fill_value)
    out = out.index_put((pos,), x)
    mask = torch.zeros(b * nmax, dtype=torch.bool, device=x.device)
    mask[pos] = True
    return out.view((b, nmax) + tuple(x.shape[1:])), mask.view(b, nmax)
