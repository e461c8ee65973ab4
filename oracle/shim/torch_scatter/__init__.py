"""ORACLE ONLY. torch_scatter 2.0.9 `scatter`/`scatter_add` restated with index_add
(reference call site kernel/go_model.py:20,200: scatter(src, index, dim=1, reduce='sum', out=zeros))."""
import torch


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce not in ("sum", "add"):
        raise NotImplementedError(reduce)
    dim = dim % src.dim()
    if out is None:
        n = int(index.max()) + 1 if dim_size is None else dim_size
        shape = list(src.shape)
        shape[dim] = n
        out = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return out.index_add(dim, index, src)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")
