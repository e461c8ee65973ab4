"""The PyTorch Geometric / torch_scatter OPERATOR surface of the path on the igcn kernels (torch_geometric 2.0.2,
torch_scatter 2.0.9 -- the versions the reference pins, environment.yml:183,211).

The reference's model files call these operators directly:

    GCNConv(in, out)(x, edge_index, edge_weight)            kernel/sgcn_img_snp.py:34-49,218-221 ; kernel/sgcn.py:25-27,281-284
    GATConv(in, out, edge_dim=1)(x, edge_index, edge_attr)   kernel/sgcn.py:163-166
    to_dense_batch(x, batch, fill_value)                     kernel/sgcn_img_snp.py:226,265,294
    global_mean_pool / global_max_pool / global_add_pool     kernel/sgcn_img_snp.py:231-233,248-250
    scatter(src, index, dim, out, reduce)                    kernel/go_model.py:20,200

Same names, same signatures, same results, differentiable with respect to x, edge_weight / edge_attr and the parameters, so the
UNMODIFIED model files run on them when `integration/shims` shadows the `torch_geometric` / `torch_scatter` import names.
(The drop-in MODEL classes -- img_snp_model.py, sgcn_models.py, go_net.py -- do not go through here: they fuse the whole
encoder into one kernel.)  CUDA tensors only; there is no CPU fallback.

  * GCNConv: any graph.  CSR by target / source is built on the device once per edge_index (cached), then the generic kernels of
    csrc/gcn_generic.cu: self-loop merge + symmetric normalisation, X W^T, warp-per-row SpMM; the backward is the transposed SpMM
    plus the normalisation's gradient into edge_weight.
  * GATConv(heads=1, edge_dim=1): the fused per-graph kernel (csrc/gat.cu).  The graph structure (equally sized graphs, one CTA
    each) is taken from the batch registry that igcn_b200.data.Batch fills; an unregistered edge_index is treated as ONE graph.
  * to_dense_batch / global_*_pool: views / fixed-order reductions when the batch vector is registered (equal sizes), the general
    torch formulation otherwise.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
from torch import nn
from torch.nn.parameter import Parameter

from . import _lib, ops
from .data import Batch, lookup_structure


# ---- graph structure cache --------------------------------------------------------------------------------------------------
class _CSR(object):
    __slots__ = ("rowptr_t", "csr_src", "csr_perm", "rowptr_s", "csc_pos", "N", "E", "keep")


_csr_cache: "OrderedDict[tuple, _CSR]" = OrderedDict()


def graph_csr(edge_index: torch.Tensor, num_nodes: int) -> _CSR:
    """Target-sorted CSR + source-sorted transposed index of an arbitrary (2, E) int64 edge_index, built on the device
    (igcn_graph_csr) and cached per (storage address, version, shape).  An entry holds a reference to the edge_index it was built from:
    while it is cached its storage cannot be freed, so the allocator cannot hand the same address to a NEW edge_index of the same
    shape (every batch of a fixed-size loader) and have it hit a stale entry; in-place edits change `_version`.  At most 8 entries."""
    _lib.require_cuda(edge_index)
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise RuntimeError("edge_index must be an int64 tensor of shape (2, E)")
    ei = edge_index.contiguous()
    key = (ei.data_ptr(), ei._version, int(ei.shape[1]), int(num_nodes), str(ei.device))
    c = _csr_cache.get(key)
    if c is not None:
        _csr_cache.move_to_end(key)
        return c
    N, E = int(num_nodes), int(ei.shape[1])
    i32 = dict(dtype=torch.int32, device=ei.device)
    c = _CSR()
    c.N, c.E = N, E
    c.rowptr_t, c.rowptr_s = torch.empty(N + 1, **i32), torch.empty(N + 1, **i32)
    c.csr_src, c.csr_perm, c.csc_pos = torch.empty(E, **i32), torch.empty(E, **i32), torch.empty(E, **i32)
    nwork = int(_lib.lib().igcn_graph_csr_work_ints(N, E))
    work = torch.empty(nwork, **i32)
    with torch.cuda.device(ei.device):
        _lib.call("igcn_graph_csr", _lib.ptr(ei), N, E, _lib.ptr(c.rowptr_t), _lib.ptr(c.csr_src), _lib.ptr(c.csr_perm), _lib.ptr(c.rowptr_s),
                  _lib.ptr(c.csc_pos), _lib.ptr(work), _lib.stream(), tag="graph_csr")
    if not torch.cuda.is_current_stream_capturing() and int(work[-1]) != 0:      # one host read per NEW edge_index
        raise RuntimeError("edge_index holds node ids outside [0, %d)" % N)
    c.keep = ei
    _csr_cache[key] = c
    while len(_csr_cache) > 8:
        _csr_cache.popitem(last=False)
    return c


# ---- GCNConv ----------------------------------------------------------------------------------------------------------------
class _GCNConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, edge_weight, weight, bias, csr: _CSR):
        _lib.require_cuda(x, edge_weight, weight, bias)
        lib = _lib.lib()
        xc, wc = x.contiguous().float(), weight.contiguous().float()
        ew = None if edge_weight is None else edge_weight.contiguous().float().view(-1)
        bc = None if bias is None else bias.contiguous().float()
        N, C = xc.shape
        O = wc.shape[0]
        if N != csr.N or (ew is not None and ew.numel() != csr.E):
            raise RuntimeError("GCNConv: x has %d rows / edge_weight %s entries, graph has %d nodes / %d edges"
                               % (N, None if ew is None else ew.numel(), csr.N, csr.E))
        saved = torch.empty(int(lib.igcn_gcn_conv_saved_floats(N, csr.E, O)), dtype=torch.float32, device=xc.device)
        out = torch.empty((N, O), dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            _lib.call("igcn_gcn_conv_fwd", _lib.ptr(xc), _lib.ptr(csr.rowptr_t), _lib.ptr(csr.csr_src), _lib.ptr(csr.csr_perm), _lib.ptr(ew),
                      _lib.ptr(wc), _lib.ptr(bc), N, csr.E, C, O, _lib.ptr(saved), _lib.ptr(out), _lib.stream(),
                      tag="gcn_conv_fwd[C=%d,O=%d]" % (C, O), nbytes=4 * (N * C + 3 * csr.E + N + 1 + N * O))
        ctx.csr, ctx.has_bias = csr, bc is not None
        ctx.need_ew = ew is not None and edge_weight.requires_grad
        ctx.save_for_backward(xc, wc, saved)
        return out

    @staticmethod
    def backward(ctx, g):
        xc, wc, saved = ctx.saved_tensors
        csr, lib = ctx.csr, _lib.lib()
        N, C = xc.shape
        O = wc.shape[0]
        g = g.contiguous().float()
        n_cta = int(lib.igcn_gcn_conv_bwd_ctas(N))
        P = O * C + O
        work = torch.empty(int(lib.igcn_gcn_conv_bwd_work_floats(N, csr.E, O)), dtype=torch.float32, device=xc.device)
        dx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        d_ew = torch.empty(csr.E, dtype=torch.float32, device=xc.device) if ctx.need_ew else None
        partials = torch.empty((n_cta, P), dtype=torch.float32, device=xc.device)
        grads = torch.empty(P, dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            _lib.call("igcn_gcn_conv_bwd", _lib.ptr(xc), _lib.ptr(csr.rowptr_t), _lib.ptr(csr.csr_src), _lib.ptr(csr.csr_perm),
                      _lib.ptr(csr.rowptr_s), _lib.ptr(csr.csc_pos), _lib.ptr(wc), _lib.ptr(saved), _lib.ptr(g), N, csr.E, C, O,
                      _lib.ptr(work), _lib.ptr(dx), _lib.ptr(d_ew), _lib.ptr(partials), n_cta, _lib.ptr(grads), _lib.stream(),
                      tag="gcn_conv_bwd[C=%d,O=%d]" % (C, O), nbytes=4 * (2 * N * C + 5 * csr.E + 2 * (N + 1) + 2 * N * O))
        return dx, d_ew, grads[:O * C].view(O, C), (grads[O * C:] if ctx.has_bias else None), None


class _Lin(nn.Module):
    """Holds `weight` so the GCN / GAT weights sit at `<conv>.lin.weight`, `<conv>.lin_src.weight` ... as in PyG 2.0.2."""

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = Parameter(torch.empty(cout, cin))

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.weight.size(0) + self.weight.size(1)))     # PyG glorot
        with torch.no_grad():
            self.weight.uniform_(-a, a)


class GCNConv(nn.Module):
    """torch_geometric.nn.GCNConv(in_channels, out_channels) with PyG 2.0.2 defaults (add_self_loops, normalize, bias).
    Parameters: `lin.weight` (out, in), `bias` (out).  forward(x, edge_index, edge_weight=None) -> (N, out)."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False, add_self_loops=True, normalize=True, bias=True, **kwargs):
        super().__init__()
        if improved or not add_self_loops or not normalize:
            raise RuntimeError("igcn_b200.pyg.GCNConv implements the default GCNConv (add_self_loops=True, normalize=True, improved=False)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels)
        self.bias = Parameter(torch.zeros(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        self.lin.reset_parameters()
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, edge_weight=None):
        csr = graph_csr(edge_index, x.shape[0])
        return _GCNConvFn.apply(x, edge_weight, self.lin.weight, self.bias, csr)

    def __repr__(self):
        return "GCNConv(%d, %d)" % (self.in_channels, self.out_channels)


# ---- GATConv ----------------------------------------------------------------------------------------------------------------
def _glorot(t):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


class GATConv(nn.Module):
    """torch_geometric.nn.GATConv(in_channels, out_channels, heads=1, edge_dim=1) as the reference uses it (kernel/sgcn.py:163-166):
    parameters lin_src / lin_dst (shared) .weight, att_src, att_dst, lin_edge.weight, att_edge, bias.
    forward(x, edge_index, edge_attr=None) -> (N, out)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0, add_self_loops=True, edge_dim=None,
                 fill_value="mean", bias=True, **kwargs):
        super().__init__()
        if heads != 1 or edge_dim != 1 or dropout != 0.0 or not add_self_loops or fill_value != "mean" or not bias:
            raise RuntimeError("igcn_b200.pyg.GATConv supports heads=1, edge_dim=1, dropout=0, add_self_loops=True, fill_value='mean' "
                               "(the reference's use) only")
        self.in_channels, self.out_channels, self.negative_slope = in_channels, out_channels, negative_slope
        self.lin_src = _Lin(in_channels, out_channels)
        self.lin_dst = self.lin_src
        self.att_src = Parameter(torch.empty(1, 1, out_channels))
        self.att_dst = Parameter(torch.empty(1, 1, out_channels))
        self.lin_edge = _Lin(1, out_channels)
        self.att_edge = Parameter(torch.empty(1, 1, out_channels))
        self.bias = Parameter(torch.zeros(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        self.lin_src.reset_parameters()
        self.lin_edge.reset_parameters()
        for t in (self.att_src, self.att_dst, self.att_edge):
            _glorot(t)
        nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, edge_attr=None, csr=None):
        if edge_attr is None:
            raise RuntimeError("igcn_b200.pyg.GATConv(edge_dim=1) needs edge_attr")
        if csr is None:
            st = lookup_structure(edge_index)
            if st is not None and st.B * st.R == x.shape[0]:
                csr = st
            else:       # an edge_index this package did not collate: one graph of N nodes (must fit one CTA)
                csr = Batch.from_device_tensors(x.detach(), edge_index, edge_attr.detach().reshape(-1), x.shape[0]).csr
        ea = edge_attr.reshape(-1).index_select(0, csr.csr_perm.long())          # per CSR slot; differentiable w.r.t. edge_attr
        return ops.gat_conv(x, csr, ea, self.lin_src.weight, self.att_src, self.att_dst, self.lin_edge.weight, self.att_edge, self.bias,
                            self.negative_slope)

    def __repr__(self):
        return "GATConv(%d, %d, heads=1)" % (self.in_channels, self.out_channels)


# ---- batch utilities ----------------------------------------------------------------------------------------------------------
def _equal_sizes(batch):
    """(B, R) when `batch` is the batch vector of a collation with equally sized graphs that this package produced, else None."""
    st = lookup_structure(batch)
    return None if st is None else (st.B, st.R)


def to_dense_batch(x, batch=None, fill_value=0, max_num_nodes=None, batch_size=None):
    """torch_geometric.utils.to_dense_batch: (N, F) node features -> ((B, N_max, F) dense, (B, N_max) bool mask)."""
    if batch is None:
        return x.unsqueeze(0), torch.ones((1, x.shape[0]), dtype=torch.bool, device=x.device)
    br = _equal_sizes(batch)
    if br is not None and br[0] * br[1] == x.shape[0] and max_num_nodes in (None, br[1]) and batch_size in (None, br[0]):
        B, R = br                                             # equally sized graphs: a view, no host synchronisation
        return x.view(B, R, *x.shape[1:]), torch.ones((B, R), dtype=torch.bool, device=x.device)
    B = int(batch.max()) + 1 if batch_size is None else int(batch_size)
    counts = torch.bincount(batch, minlength=B)
    cum = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    n_max = int(counts.max()) if max_num_nodes is None else int(max_num_nodes)
    idx = torch.arange(batch.numel(), device=x.device) - cum[batch] + batch * n_max
    out = x.new_full((B * n_max,) + tuple(x.shape[1:]), fill_value)
    out[idx] = x
    mask = torch.zeros(B * n_max, dtype=torch.bool, device=x.device)
    mask[idx] = True
    return out.view(B, n_max, *x.shape[1:]), mask.view(B, n_max)


def _pool(x, batch, size, how):
    if batch is None:
        r = {"sum": x.sum(0, keepdim=True), "mean": x.mean(0, keepdim=True), "max": x.max(0, keepdim=True)[0]}
        return r[how]
    br = _equal_sizes(batch)
    if br is not None and br[0] * br[1] == x.shape[0] and size in (None, br[0]):
        v = x.view(br[0], br[1], *x.shape[1:])               # fixed-order reduction per graph, no atomics
        return {"sum": v.sum(1), "mean": v.mean(1), "max": v.max(1)[0]}[how]
    B = int(batch.max()) + 1 if size is None else int(size)
    idx = batch.view(-1, *([1] * (x.dim() - 1))).expand_as(x)
    if how == "max":
        return x.new_full((B,) + tuple(x.shape[1:]), float("-inf")).scatter_reduce(0, idx, x, "amax", include_self=True)
    out = x.new_zeros((B,) + tuple(x.shape[1:])).index_add(0, batch, x)
    if how == "mean":
        out = out / torch.bincount(batch, minlength=B).clamp(min=1).view(-1, *([1] * (x.dim() - 1))).to(x.dtype)
    return out


def global_add_pool(x, batch, size=None):
    return _pool(x, batch, size, "sum")


def global_mean_pool(x, batch, size=None):
    return _pool(x, batch, size, "mean")


def global_max_pool(x, batch, size=None):
    return _pool(x, batch, size, "max")


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    """torch_scatter.scatter for reduce in {'sum', 'add', 'mean', 'max', 'min'} (1-D `index` along `dim`, as the reference calls it:
    kernel/go_model.py:200 `scatter(src, index, dim=1, reduce='sum')`)."""
    dim = dim if dim >= 0 else src.dim() + dim
    if index.dim() != 1:
        index = index.reshape(-1) if index.numel() == src.shape[dim] else index.select(-1, 0)
    n = int(dim_size) if dim_size is not None else (out.shape[dim] if out is not None else int(index.max()) + 1)
    shape = list(src.shape)
    shape[dim] = n
    if reduce in ("sum", "add"):
        base = src.new_zeros(shape) if out is None else out
        return base.index_add(dim, index, src) if out is None else out.index_add_(dim, index, src)
    view = [1] * src.dim()
    view[dim] = -1
    idx = index.view(view).expand_as(src)
    if reduce == "mean":
        s = src.new_zeros(shape).index_add(dim, index, src)
        c = torch.bincount(index, minlength=n).clamp(min=1).to(src.dtype).view(view)
        return s / c
    if reduce in ("max", "min"):
        init = float("-inf") if reduce == "max" else float("inf")
        return src.new_full(shape, init).scatter_reduce(dim, idx, src, "amax" if reduce == "max" else "amin", include_self=True)
    raise RuntimeError("igcn_b200.pyg.scatter: unsupported reduce=%r" % (reduce,))


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")
