"""ORACLE ONLY. torch_geometric.nn 2.0.2 restated in plain torch (CPU, any dtype).

Call sites in the reference that this serves:
  GCNConv  -- kernel/sgcn_img_snp.py:34,40,42,49 ; kernel/sgcn.py:25-27,281,284
  GATConv  -- kernel/sgcn.py:163,166
  global_{mean,max,add}_pool -- kernel/sgcn_img_snp.py:231-233,248-250
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from ..utils import (add_remaining_self_loops, add_self_loops,
                     remove_self_loops, segment_softmax)


def _glorot(t):
    # PyG `glorot`: U(-a, a), a = sqrt(6 / (fan_in + fan_out)) on the last two dims
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


class _PygLinear(nn.Module):
    """PyG's own `Linear(in, out, bias=False, weight_initializer='glorot')`.
    The GCN weight therefore lives at state_dict key `<conv>.lin.weight`."""

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.reset_parameters()

    def reset_parameters(self):
        _glorot(self.weight)

    def forward(self, x):
        return F.linear(x, self.weight)


def gcn_norm(edge_index, edge_weight, num_nodes, improved=False, self_loops=True, dtype=None):
    """PyG 2.0.2 gcn_norm: self-loop merge (fill 1, or 2 if improved), in-degree
    by TARGET (edge_index[1]), symmetric D^-1/2 A D^-1/2, inf -> 0."""
    fill = 2.0 if improved else 1.0
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    if self_loops:
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill, num_nodes)
    src, dst = edge_index[0], edge_index[1]
    deg = torch.zeros(num_nodes, dtype=edge_weight.dtype, device=edge_weight.device)
    deg = deg.index_add(0, dst, edge_weight)
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0.0)
    return edge_index, dis[src] * edge_weight * dis[dst]


class GCNConv(nn.Module):
    def __init__(self, in_channels, out_channels, improved=False, cached=False,
                 add_self_loops=True, normalize=True, bias=True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.add_self_loops, self.normalize = improved, add_self_loops, normalize
        self.lin = _PygLinear(in_channels, out_channels)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def reset_parameters(self):
        self.lin.reset_parameters()
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, edge_weight=None):
        n = x.size(0)
        if self.normalize:
            edge_index, edge_weight = gcn_norm(edge_index, edge_weight, n, self.improved,
                                               self.add_self_loops, x.dtype)
        h = self.lin(x)
        msg = h[edge_index[0]]
        if edge_weight is not None:
            msg = edge_weight.view(-1, 1) * msg
        # aggr='add' at the target, in edge order (self loops were appended last)
        out = torch.zeros(n, h.size(1), dtype=h.dtype, device=h.device).index_add(0, edge_index[1], msg)
        if self.bias is not None:
            out = out + self.bias
        return out


class GATConv(nn.Module):
    """PyG 2.0.2 GATConv (heads, concat, edge_dim, fill_value='mean')."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0.0, add_self_loops=True, edge_dim=None, fill_value="mean", bias=True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops, self.edge_dim, self.fill_value = add_self_loops, edge_dim, fill_value
        self.lin_src = _PygLinear(in_channels, heads * out_channels)
        self.lin_dst = self.lin_src
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        if edge_dim is not None:
            self.lin_edge = _PygLinear(edge_dim, heads * out_channels)
            self.att_edge = nn.Parameter(torch.empty(1, heads, out_channels))
        else:
            self.lin_edge = None
            self.register_parameter("att_edge", None)
        if bias:
            self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        self.lin_src.reset_parameters()
        if self.lin_edge is not None:
            self.lin_edge.reset_parameters()
            _glorot(self.att_edge)
        _glorot(self.att_src)
        _glorot(self.att_dst)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, edge_attr=None):
        H, C, n = self.heads, self.out_channels, x.size(0)
        h = self.lin_src(x).view(-1, H, C)
        a_src = (h * self.att_src).sum(-1)
        a_dst = (h * self.att_dst).sum(-1)
        if self.add_self_loops:
            edge_index, edge_attr = remove_self_loops(edge_index, edge_attr)
            edge_index, edge_attr = add_self_loops(edge_index, edge_attr, self.fill_value, n)
        src, dst = edge_index[0], edge_index[1]
        alpha = a_src[src] + a_dst[dst]
        if edge_attr is not None and self.lin_edge is not None:
            ea = edge_attr.view(-1, 1) if edge_attr.dim() == 1 else edge_attr
            ea = self.lin_edge(ea).view(-1, H, C)
            alpha = alpha + (ea * self.att_edge).sum(-1)
        alpha = F.leaky_relu(alpha, self.negative_slope)
        alpha = segment_softmax(alpha, dst, n)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        msg = h[src] * alpha.unsqueeze(-1)
        out = torch.zeros(n, H, C, dtype=h.dtype, device=h.device).index_add(0, dst, msg)
        out = out.view(-1, H * C) if self.concat else out.mean(1)
        if self.bias is not None:
            out = out + self.bias
        return out


class ChebConv(nn.Module):  # imported by the reference, never constructed on the hot path
    def __init__(self, *a, **k):
        raise NotImplementedError("ChebConv is outside the IG-GCN hot path")


def _num_graphs(batch, size):
    return int(batch.max().item()) + 1 if size is None else size


def global_add_pool(x, batch, size=None):
    b = _num_graphs(batch, size)
    return torch.zeros(b, x.size(1), dtype=x.dtype, device=x.device).index_add(0, batch, x)


def global_mean_pool(x, batch, size=None):
    b = _num_graphs(batch, size)
    cnt = torch.zeros(b, dtype=x.dtype, device=x.device).index_add(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return global_add_pool(x, batch, b) / cnt.clamp(min=1).view(-1, 1)


def global_max_pool(x, batch, size=None):
    b = _num_graphs(batch, size)
    out = torch.full((b, x.size(1)), float("-inf"), dtype=x.dtype, device=x.device)
    return out.scatter_reduce(0, batch.view(-1, 1).expand_as(x), x, reduce="amax", include_self=True)


def global_sort_pool(*a, **k):
    raise NotImplementedError("global_sort_pool is outside the IG-GCN hot path")
