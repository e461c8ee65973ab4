"""autograd wrappers around the C-ABI kernels (include/igcn_b200.h).  CUDA tensors only."""
from __future__ import annotations

import torch

import os

from . import _lib
from .data import GraphCSR


def pack_layer_params(weights, biases):
    """[W_1 (H,F0) | b_1 | W_2 (H,H) | b_2 | ...] -- the `wb` layout of igcn_sgcn_encoder_*."""
    parts = []
    for w, b in zip(weights, biases):
        parts.append(w.reshape(-1))
        parts.append(b.reshape(-1))
    return torch.cat(parts) if parts else None


def _alias(t):
    """A new tensor over the storage of the contiguous tensor t that autograd does not track as a view of it."""
    return torch.empty(0, dtype=t.dtype, device=t.device).set_(t.untyped_storage(), t.storage_offset(), t.shape, t.stride())


def _adjacent_view(tensors):
    """If the tensors sit back to back in one storage (train.FlatAdam lays the parameters out in registration order, which for
    the SGCN layers IS the wb order), return a zero-copy 1-D view over all of them; else None."""
    if not tensors:
        return None
    t0 = tensors[0]
    if not t0.is_contiguous():
        return None
    ptr, n = t0.data_ptr(), 0
    for t in tensors:
        if not t.is_contiguous() or t.dtype != torch.float32 or t.data_ptr() != ptr + 4 * n or t.device != t0.device:
            return None
        n += t.numel()
    try:
        return torch.as_strided(t0.detach(), (n,), (1,), t0.storage_offset())
    except RuntimeError:
        return None


class _SplitGradFn(torch.autograd.Function):
    """wb = concatenation of the layer parameters WITHOUT a copy when they are adjacent in memory (one cat kernel per encoder call
    and its split in the backward disappear; the gradients come back as views of the kernel's gradient buffer)."""

    @staticmethod
    def forward(ctx, *tensors):
        ctx.shapes = [t.shape for t in tensors]
        v = _adjacent_view(list(tensors))
        if v is None:
            v = torch.cat([t.reshape(-1) for t in tensors])
        return v

    @staticmethod
    def backward(ctx, g):
        out, off = [], 0
        for sh in ctx.shapes:
            n = 1
            for d in sh:
                n *= d
            out.append(g[off:off + n].view(sh))
            off += n
        return tuple(out)


class _SGCNEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, prob, prob_bias, wb, csr: GraphCSR, L: int, H: int, want_pe: bool, relu: bool = True, out_buf=None):
        _lib.require_cuda(x, prob, prob_bias, wb)
        lib = _lib.lib()
        x = x.contiguous().float()
        B, R, F0 = csr.B, csr.R, x.shape[1]
        if x.shape[0] != B * R:
            raise RuntimeError("sgcn_encoder: x has %d rows, batch structure has %d" % (x.shape[0], B * R))
        explain = prob is not None
        probc = prob.contiguous().float() if explain else None
        pbc = prob_bias.contiguous().float().view(-1) if explain else None
        wbc = wb.contiguous().float() if wb is not None else None
        if out_buf is not None:                  # caller-provided destination (a slice of the stacked two-pass buffer): no copy later
            if tuple(out_buf.shape) != (B, R, L * H) or not out_buf.is_contiguous() or out_buf.dtype != torch.float32:
                raise RuntimeError("sgcn_encoder: out buffer must be a contiguous float32 (%d, %d, %d) tensor" % (B, R, L * H))
            out = _alias(out_buf)                # same storage, no view relationship (the buffer has no autograd history)
        else:
            out = torch.empty((B, R, L * H), dtype=torch.float32, device=x.device)
        p_e = torch.empty(csr.E, dtype=torch.float32, device=x.device) if (explain and want_pe) else None
        with torch.cuda.device(x.device):
            _lib.call("igcn_sgcn_encoder_fwd", _lib.ptr(x), _lib.ptr(csr.rowptr_t), _lib.ptr(csr.csr_src), _lib.ptr(csr.csr_w),
                                           _lib.ptr(probc), _lib.ptr(pbc), _lib.ptr(wbc), B, R, F0, H, L, csr.max_eg, int(relu),
                                           _lib.ptr(out), _lib.ptr(p_e), _lib.stream(),
                      tag="sgcn_encoder_fwd[%s,L=%d]" % ("explain" if explain else "plain", L),
                      nbytes=4 * (B * R * F0 + 2 * csr.E + B * R + 1 + B * R * L * H + (csr.E + R * F0 if explain else 0)))
        ctx.csr, ctx.L, ctx.H, ctx.explain, ctx.relu = csr, L, H, explain, bool(relu)
        ctx.save_for_backward(x, probc, pbc, wbc, out)
        if p_e is None:
            p_e = x.new_empty(0)
            ctx.mark_non_differentiable(p_e)
        return out, p_e

    @staticmethod
    def backward(ctx, g_out, g_pe):
        x, prob, pb, wb, out = ctx.saved_tensors
        csr, L, H = ctx.csr, ctx.L, ctx.H
        lib = _lib.lib()
        B, R, F0 = csr.B, csr.R, x.shape[1]
        P = lib.igcn_sgcn_param_count(R, F0, H, L)
        n_cta = lib.igcn_sgcn_bwd_ctas(B, R, F0, H, L, csr.max_eg)
        dx = torch.empty_like(x)
        partials = torch.empty((max(n_cta, 1), P), dtype=torch.float32, device=x.device)
        grads = torch.empty(P, dtype=torch.float32, device=x.device)
        g_out = g_out.contiguous() if L > 0 else None
        gpe = g_pe.contiguous() if (ctx.explain and g_pe is not None and g_pe.numel() == csr.E and csr.E > 0) else None
        with torch.cuda.device(x.device):
            _lib.call("igcn_sgcn_encoder_bwd", _lib.ptr(x), _lib.ptr(csr.rowptr_t), _lib.ptr(csr.csr_src), _lib.ptr(csr.csr_w),
                                           _lib.ptr(csr.rowptr_s), _lib.ptr(csr.csc_pos), _lib.ptr(prob), _lib.ptr(pb),
                                           _lib.ptr(wb), _lib.ptr(out), _lib.ptr(g_out), _lib.ptr(gpe), B, R, F0, H, L,
                                           csr.max_eg, int(ctx.relu), _lib.ptr(dx), _lib.ptr(partials), n_cta, _lib.ptr(grads), _lib.stream(),
                      tag="sgcn_encoder_bwd[%s,L=%d]" % ("explain" if ctx.explain else "plain", L),
                      nbytes=4 * (2 * B * R * L * H + 2 * B * R * F0 + 3 * csr.E + 2 * (B * R + 1) + (csr.E if gpe is not None else 0)))
        nwb = P - R * F0 - 2 * F0
        d_wb = grads[:nwb] if wb is not None else None
        d_prob = grads[nwb:nwb + R * F0].view(R, F0) if ctx.explain else None
        d_pb = grads[nwb + R * F0:].view(2 * F0, 1) if ctx.explain else None
        return dx, d_prob, d_pb, d_wb, None, None, None, None, None, None


class _JoinHalvesFn(torch.autograd.Function):
    """The stacked (2B, ...) tensor whose halves ARE `a` and `b` (two adjacent slices of one buffer, each written in place by its
    producer): no copy forward, two views backward.  Replaces torch.cat([a, b], 0) of kernel-produced halves."""

    @staticmethod
    def forward(ctx, a, b, whole):
        if a.data_ptr() != whole.data_ptr() or b.data_ptr() != whole.data_ptr() + a.numel() * a.element_size() \
                or a.numel() + b.numel() != whole.numel():
            raise RuntimeError("join_halves: the halves are not the two halves of the buffer")
        ctx.n = a.shape[0]
        return _alias(whole)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        return g[:ctx.n], g[ctx.n:], None


def sgcn_encoder(x, csr: GraphCSR, weights, biases, prob=None, prob_bias=None, want_pe=False, relu=True, out=None):
    """Fused SGCN encoder. Returns (out (B,R,L*H), p_e (E,) in CSR-slot order or empty).  `out`: optional destination buffer (a
    contiguous (B,R,L*H) float32 tensor without autograd history, e.g. one half of a stacked two-pass buffer; see join_halves)."""
    L = len(weights)
    H = weights[0].shape[0] if L else 0
    wb = _SplitGradFn.apply(*[t for w, b in zip(weights, biases) for t in (w, b)]) if L else None
    return _SGCNEncoderFn.apply(x, prob, prob_bias, wb, csr, L, H, want_pe, relu, out)


class _FanOutFn(torch.autograd.Function):
    """k aliases of one tensor for k consumers.  Backward: ONE sum of the consumers' gradients, taken when the last has arrived.
    autograd's own accumulation adds the gradients pairwise as they arrive, each add on the producer's stream -- which puts the
    fast consumers' backward chains in front of the slow consumer's on that stream (device timeline)."""

    @staticmethod
    def forward(ctx, x, k):
        x = x.contiguous()
        return tuple(_alias(x) for _ in range(k))

    @staticmethod
    def backward(ctx, *gs):
        gs = [g.contiguous() for g in gs if g is not None]
        if not gs:
            return None, None
        if len(gs) == 1:
            return gs[0], None
        acc = gs[0]
        rest = gs[1:]
        while rest:
            b, c = rest[0], (rest[1] if len(rest) > 1 else None)
            rest = rest[2:]
            out = torch.empty_like(acc)
            with torch.cuda.device(acc.device):
                _lib.call("igcn_sum3", _lib.ptr(acc), _lib.ptr(b), _lib.ptr(c), acc.numel(), _lib.ptr(out), _lib.stream(), tag="fan_out_sum",
                          nbytes=4 * acc.numel() * (3 + (c is not None)))
            acc = out
        return acc, None


def fan_out(x, k):
    """x for k consumers whose gradients are summed in one launch (CUDA tensors that need a gradient; otherwise x itself k times)."""
    if k < 2 or not x.is_cuda or not (torch.is_grad_enabled() and x.requires_grad) or os.environ.get("IGCN_NO_FAN_OUT") == "1":
        return (x,) * k
    return _FanOutFn.apply(x, k)


def join_halves(a, b, whole):
    """`whole` (2B, ...) as an autograd tensor whose halves are a and b -- the tensors two producers wrote into whole[:B] / whole[B:]
    through their `out=` argument -- without the copy of torch.cat([a, b], 0)."""
    return _JoinHalvesFn.apply(a, b, whole)


def edge_mask(x, csr: GraphCSR, prob, prob_bias):
    """p_e of cal_probability (kernel/sgcn_img_snp.py:141-142) in CSR-slot order (masks only, L=0)."""
    _, p_e = _SGCNEncoderFn.apply(x, prob, prob_bias, None, csr, 0, 0, True, True)
    return p_e


class _GATConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ea_csr, W, att_src, att_dst, lin_edge, att_edge, bias, csr: GraphCSR, slope: float):
        _lib.require_cuda(x, ea_csr, W)
        c = lambda t: t.contiguous().float().view(-1)
        x = x.contiguous().float()
        B, R, Fin, H = csr.B, csr.R, x.shape[1], W.shape[0]
        Wc, a_s, a_d, le, ae, b = W.contiguous().float(), c(att_src), c(att_dst), c(lin_edge), c(att_edge), c(bias)
        ea = c(ea_csr)
        out = torch.empty((B * R, H), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("igcn_gat_layer_fwd", _lib.ptr(x), _lib.ptr(csr.rowptr_t), _lib.ptr(csr.csr_src), _lib.ptr(ea), _lib.ptr(Wc),
                      _lib.ptr(a_s), _lib.ptr(a_d), _lib.ptr(le), _lib.ptr(ae), _lib.ptr(b), B, R, Fin, H, csr.max_eg, float(slope),
                      _lib.ptr(out), _lib.stream(), tag="gat_layer_fwd[Fin=%d,H=%d]" % (Fin, H),
                      nbytes=4 * (B * R * Fin + 3 * csr.E + B * R + 1 + B * R * H))
        ctx.csr, ctx.slope = csr, float(slope)
        ctx.shapes = (att_src.shape, att_dst.shape, lin_edge.shape, att_edge.shape, bias.shape)
        ctx.save_for_backward(x, ea, Wc, a_s, a_d, le, ae, b)
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, ea, W, a_s, a_d, le, ae, b = ctx.saved_tensors
        csr, lib = ctx.csr, _lib.lib()
        B, R, Fin, H = csr.B, csr.R, x.shape[1], W.shape[0]
        P = lib.igcn_gat_param_count(Fin, H)
        n_cta = lib.igcn_gat_bwd_ctas(B, R, Fin, H, csr.max_eg)
        dx, d_ea = torch.empty_like(x), torch.empty_like(ea)
        partials = torch.empty((max(n_cta, 1), P), dtype=torch.float32, device=x.device)
        grads = torch.empty(P, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("igcn_gat_layer_bwd", _lib.ptr(x), _lib.ptr(csr.rowptr_t), _lib.ptr(csr.csr_src), _lib.ptr(ea),
                      _lib.ptr(csr.rowptr_s), _lib.ptr(csr.csc_pos), _lib.ptr(W), _lib.ptr(a_s), _lib.ptr(a_d), _lib.ptr(le), _lib.ptr(ae),
                      _lib.ptr(b), _lib.ptr(g_out.contiguous()), B, R, Fin, H, csr.max_eg, ctx.slope, _lib.ptr(dx), _lib.ptr(d_ea),
                      _lib.ptr(partials), n_cta, _lib.ptr(grads), _lib.stream(), tag="gat_layer_bwd[Fin=%d,H=%d]" % (Fin, H),
                      nbytes=4 * (2 * B * R * Fin + 5 * csr.E + 2 * (B * R + 1) + B * R * H))
        o = H * Fin
        sh = ctx.shapes
        return (dx, d_ea, grads[:o].view(H, Fin), grads[o:o + H].view(sh[0]), grads[o + H:o + 2 * H].view(sh[1]),
                grads[o + 2 * H:o + 3 * H].view(sh[2]), grads[o + 3 * H:o + 4 * H].view(sh[3]), grads[o + 4 * H:o + 5 * H].view(sh[4]),
                None, None)


def gat_conv(x, csr: GraphCSR, edge_attr_csr, W, att_src, att_dst, lin_edge, att_edge, bias, negative_slope=0.2):
    """Fused GATConv(heads=1, edge_dim=1). edge_attr_csr is per CSR slot (edge_attr[csr.csr_perm])."""
    return _GATConvFn.apply(x, edge_attr_csr, W, att_src, att_dst, lin_edge, att_edge, bias, csr, negative_slope)


# ---- tensor-core (tcgen05, 3xTF32) products: ops built on igcn_tc_split + igcn_tc_gemm ---------------------------------
USE_TC = os.environ.get("IGCN_NO_TC", "") == ""     # IGCN_NO_TC=1 keeps the fp32 FFMA tile kernels (A/B comparison)


def _pad4(n):
    return (n + 3) // 4 * 4


_AUX_STREAM = os.environ.get("IGCN_NO_AUX_STREAM", "") == ""
_aux_streams = {}


def _aux_stream(dev, which=0):
    st = _aux_streams.get((str(dev), which))
    if st is None:
        st = torch.cuda.Stream(device=dev)
        _aux_streams[(str(dev), which)] = st
    return st


def _job(src, hi_lo, rows, cols, ld_src, row_off=0, col_off=0, transpose=False, mask=None, sub=None):
    """One operand-preparation job of igcn_tc_split; hi_lo is a (2, R, ld) buffer (hi = [0], lo = [1])."""
    return [0 if src is None else src.data_ptr(), 0 if mask is None else mask.data_ptr(), hi_lo[0].data_ptr(), hi_lo[1].data_ptr(),
            rows, cols, ld_src, hi_lo.shape[2], row_off, col_off, int(transpose), 0 if sub is None else sub.data_ptr()]


def _tc_split(jobs, dev):
    import ctypes
    flat = [v for j in jobs for v in j]
    arr = (ctypes.c_int64 * len(flat))(*flat)
    with torch.cuda.device(dev):
        _lib.call("igcn_tc_split", ctypes.addressof(arr), len(jobs), _lib.stream(), tag="tc_split[%d jobs]" % len(jobs))


def _tc_gemm(a, b, M, N, K, dsts, widths, strides, bias=None, relu=False, tag="tc_gemm"):
    """C[m][n] = act(sum_k A[m][k] B[n][k] + bias[n]); a, b: (2, rows, ld) hi/lo buffers; C columns go to `dsts`."""
    import ctypes
    lib = _lib.lib()
    S = lib.igcn_tc_gemm_splits(M, N, K)
    part = torch.empty((S, M, N), dtype=torch.float32, device=a.device) if S > 1 else None
    dsts = list(dsts) + [None] * (3 - len(dsts))
    hw = (ctypes.c_int64 * 3)(*(list(widths) + [0] * (3 - len(widths))))
    hs = (ctypes.c_int64 * 3)(*(list(strides) + [0] * (3 - len(strides))))
    with torch.cuda.device(a.device):
        _lib.call("igcn_tc_gemm", a[0].data_ptr(), a[1].data_ptr(), a.shape[2], b[0].data_ptr(), b[1].data_ptr(), b.shape[2], M, N, K,
                  _lib.ptr(bias), int(relu), _lib.ptr(dsts[0]), _lib.ptr(dsts[1]), _lib.ptr(dsts[2]), ctypes.addressof(hw),
                  ctypes.addressof(hs), _lib.ptr(part), S, _lib.stream(), tag="%s[M=%d,N=%d,K=%d]" % (tag, M, N, K),
                  nbytes=4 * (2 * M * K + 2 * N * K + M * N))


def tc_matmul_nt(a, b, bias=None, relu=False):
    """act(a @ b.T + bias) for CUDA fp32 matrices a (M,K), b (N,K) on the tcgen05 tensor cores with fp32-level accuracy
    (3xTF32).  No autograd; the building block of cat_linear / laplacian_quadratic, exposed for tests."""
    _lib.require_cuda(a, b, bias)
    M, K = a.shape
    N = b.shape[0]
    a, b = a.contiguous().float(), b.contiguous().float()
    A = torch.empty((2, M, _pad4(K)), dtype=torch.float32, device=a.device)
    Bm = torch.empty((2, N, _pad4(K)), dtype=torch.float32, device=a.device)
    _tc_split([_job(a, A, M, K, K), _job(b, Bm, N, K, K)], a.device)
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _tc_gemm(A, Bm, M, N, K, [out], [N], [N], bias, relu)
    return out


# ---- operands shared by several heads: ONE split launch ------------------------------------------------------------------------------
_PRESPLIT = {}


def _src_key(t):
    return None if t is None else (t.data_ptr(), tuple(t.shape), t.stride(0))


def presplit_heads(xs, weights):
    """The fusion heads lin1 / lin1_regr read the same concatenation (lin1 a column prefix of it): their tcgen05 operands are prepared
    by ONE igcn_tc_split launch here -- the (hi, lo) image of [x0 | x1 | x2] and of every weight -- and the cat_linear calls that follow
    in the same forward pick them up (one shot: every entry is consumed by the call that uses it, so nothing stale survives a
    parameter update).  No-op when the tensor-core path is off or a weight does not fit."""
    _PRESPLIT.clear()
    xs = list(xs) + [None] * (3 - len(xs))
    if not USE_TC or any(t is not None and not t.is_cuda for t in xs):
        return
    M = max(t.shape[0] for t in xs if t is not None)
    reps = [1 if t is None else M // t.shape[0] for t in xs]
    if any(t is not None and t.shape[0] * r != M for t, r in zip(xs, reps)) or M == 0:
        return
    cs = [None if t is None else (t.detach().float() if t.stride(-1) == 1 else t.detach().contiguous().float()) for t in xs]
    widths = [0 if t is None else t.shape[1] for t in cs]
    K = sum(widths)
    dev = next(t for t in cs if t is not None).device
    A = torch.empty((2, M, _pad4(K)), dtype=torch.float32, device=dev)
    jobs, off = [], 0
    for t, w, rp in zip(cs, widths, reps):
        if t is not None and w:
            for k in range(rp):
                jobs.append(_job(t, A, M // rp, w, t.stride(0), row_off=k * (M // rp), col_off=off))
        off += w
    wsplit = {}
    for W in weights:
        Wc = W.detach().contiguous().float()
        if Wc.shape[1] > K:
            _PRESPLIT.clear()
            return
        Bw = torch.empty((2, Wc.shape[0], _pad4(Wc.shape[1])), dtype=torch.float32, device=dev)
        jobs.append(_job(Wc, Bw, Wc.shape[0], Wc.shape[1], Wc.shape[1]))
        wsplit[Wc.data_ptr()] = Bw
    if len(jobs) > 8:
        return
    _tc_split(jobs, dev)
    _PRESPLIT.update(A=A, M=M, srcs=[_src_key(t) for t in cs], reps=reps, widths=widths, weights=wsplit)


def _presplit_take(cs, widths, reps, Wc, M):
    """(A, Bw) prepared by presplit_heads for exactly these sources (a prefix of the prepared concatenation) and this weight."""
    c = _PRESPLIT
    if not c or c["M"] != M or Wc.data_ptr() not in c["weights"]:
        return None
    n = max(i + 1 for i, t in enumerate(cs) if t is not None)
    for i in range(3):
        if i < n:
            if _src_key(cs[i]) != c["srcs"][i] or reps[i] != c["reps"][i] or widths[i] != c["widths"][i]:
                return None
        elif cs[i] is not None:
            return None
    return c["A"], c["weights"].pop(Wc.data_ptr())


# Weight gradients computed on an auxiliary stream are read only by the optimizer: inside train.train_step (which opens the scope,
# and joins before it gathers the gradients) the consumer stream does not wait for them where they are produced.
_DEFERRED = None            # None, or {"events": [(device, event)], "seen": set(id(weight))}


def defer_weight_grad_joins():
    global _DEFERRED
    _DEFERRED = dict(events=[], seen=set())


def join_weight_grads():
    """Makes the CURRENT stream (the one the optimizer runs on) wait for the deferred weight-gradient products; closes the scope."""
    global _DEFERRED
    d, _DEFERRED = _DEFERRED, None
    if d:
        for dev, ev in d["events"]:
            torch.cuda.current_stream(dev).wait_event(ev)


def _join_or_defer(cur, aux, W):
    d = _DEFERRED
    if d is None or os.environ.get("IGCN_NO_DEFER_DW") == "1":
        cur.wait_stream(aux)
        return
    if id(W) in d["seen"]:
        # a second use of the same weight: autograd will ADD this gradient to the earlier one on `cur`, so both must have landed
        for _, ev in d["events"]:
            cur.wait_event(ev)
        cur.wait_stream(aux)
        return
    d["seen"].add(id(W))
    ev = torch.cuda.Event()
    ev.record(aux)
    d["events"].append((cur.device, ev))


class _CatLinearFn(torch.autograd.Function):
    """act([x0 | x1 | x2] W^T + b) via igcn_cat_linear_* (sources may be None)."""

    @staticmethod
    def forward(ctx, x0, x1, x2, W, bias, relu: bool):
        import ctypes
        srcs = [x0, x1, x2]
        _lib.require_cuda(W, bias, *[t for t in srcs if t is not None])
        lib = _lib.lib()
        M = max(t.shape[0] for t in srcs if t is not None)
        # a source with M / k rows is read k times (rows i, i + M/k, ...): the stacked plain / explain passes share `img_feat`
        reps = [1 if t is None else M // t.shape[0] for t in srcs]
        if any(t is not None and t.shape[0] * r != M for t, r in zip(srcs, reps)):
            raise RuntimeError("cat_linear: source row counts %s do not divide %d" % ([None if t is None else t.shape[0] for t in srcs], M))
        if not (USE_TC and M > 0) and any(r > 1 for r in reps):
            srcs = [t if r == 1 else torch.cat([t] * r, 0) for t, r in zip(srcs, reps)]
            reps = [1, 1, 1]
        ctx.reps = reps
        cs = [None if t is None else (t.float() if t.stride(-1) == 1 else t.contiguous().float()) for t in srcs]
        widths = [0 if t is None else t.shape[1] for t in cs]
        strides = [0 if t is None else t.stride(0) for t in cs]
        N, K = W.shape
        if sum(widths) != K:
            raise RuntimeError("cat_linear: source widths %s do not add up to in_features=%d" % (widths, K))
        Wc, bc = W.contiguous().float(), bias.contiguous().float()
        ctx.set_materialize_grads(False)
        # small batches: warp-level tensor-core kernels that read the sources in place (csrc/catlin_mma.cu), one launch per product
        ctx.mma = bool(USE_TC and M > 0 and lib.igcn_catlin_mma_supported(M, N, K))
        if ctx.mma:
            out = torch.empty((M, N), dtype=torch.float32, device=W.device)
            rows = [1 if t is None else t.shape[0] for t in cs]
            hw, hs, hr = (ctypes.c_int64 * 3)(*widths), (ctypes.c_int64 * 3)(*strides), (ctypes.c_int64 * 3)(*rows)
            with torch.cuda.device(W.device):
                _lib.call("igcn_catlin_mma_fwd", _lib_ptr_strided(cs[0]), _lib_ptr_strided(cs[1]), _lib_ptr_strided(cs[2]),
                          ctypes.addressof(hw), ctypes.addressof(hs), ctypes.addressof(hr), _lib.ptr(Wc), _lib.ptr(bc), M, N, K, int(relu),
                          _lib.ptr(out), _lib.stream(), tag="cat_linear_fwd_mma[M=%d,N=%d,K=%d]" % (M, N, K), nbytes=4 * (M * K + N * K + M * N))
            ctx.relu, ctx.widths, ctx.strides, ctx.rows, ctx.tc = bool(relu), widths, strides, rows, False
            ctx.need = [t is not None and t.requires_grad for t in srcs]
            ctx.save_for_backward(Wc, out, *[t for t in cs if t is not None])
            ctx.present = [t is not None for t in cs]
            return out
        ctx.tc = USE_TC and M > 0
        if ctx.tc:
            # tensor cores: one split launch folds the concatenation, then one 3xTF32 product with bias + ReLU in its epilogue
            pre = _presplit_take(cs, widths, reps, Wc, M)
            if pre is not None:                  # operands already prepared by presplit_heads (shared with the other head)
                A, Bw = pre
            else:
                A = torch.empty((2, M, _pad4(K)), dtype=torch.float32, device=W.device)
                Bw = torch.empty((2, N, _pad4(K)), dtype=torch.float32, device=W.device)
                jobs, off = [], 0
                for t, w, ld, rp in zip(cs, widths, strides, reps):
                    if t is not None and w:
                        for k in range(rp):
                            jobs.append(_job(t, A, M // rp, w, ld, row_off=k * (M // rp), col_off=off))
                    off += w
                jobs.append(_job(Wc, Bw, N, K, K))
                _tc_split(jobs, W.device)
            out = torch.empty((M, N), dtype=torch.float32, device=W.device)
            _tc_gemm(A, Bw, M, N, K, [out], [N], [N], bc, relu, tag="cat_linear_fwd_tc")
            ctx.relu, ctx.widths, ctx.strides = bool(relu), widths, strides
            ctx.need = [t is not None and t.requires_grad for t in srcs]
            ctx.save_for_backward(Wc, out, *[t for t in cs if t is not None])
            ctx.present = [t is not None for t in cs]
            return out
        S = lib.igcn_cat_linear_splits(M, N, K)
        part = torch.empty((S, M, N), dtype=torch.float32, device=W.device)
        out = torch.empty((M, N), dtype=torch.float32, device=W.device)
        hw, hs = (ctypes.c_int64 * 3)(*widths), (ctypes.c_int64 * 3)(*strides)
        with torch.cuda.device(W.device):
            _lib.call("igcn_cat_linear_fwd", _lib_ptr_strided(cs[0]), _lib_ptr_strided(cs[1]), _lib_ptr_strided(cs[2]),
                      ctypes.addressof(hw), ctypes.addressof(hs), _lib.ptr(Wc), _lib.ptr(bc), M, N, K, int(relu), _lib.ptr(part), S,
                      _lib.ptr(out), _lib.stream(), tag="cat_linear_fwd[M=%d,N=%d,K=%d]" % (M, N, K), nbytes=4 * (M * K + N * K + M * N))
        ctx.relu, ctx.widths, ctx.strides = bool(relu), widths, strides
        ctx.need = [t is not None and t.requires_grad for t in srcs]
        ctx.save_for_backward(Wc, out, *[t for t in cs if t is not None])
        ctx.present = [t is not None for t in cs]
        return out

    @staticmethod
    def backward(ctx, g_out):
        import ctypes
        if g_out is None:                      # this head does not reach the loss
            return None, None, None, None, None, None
        saved = ctx.saved_tensors
        W, out = saved[0], saved[1]
        it = iter(saved[2:])
        cs = [next(it) if p else None for p in ctx.present]
        M, (N, K) = out.shape[0], W.shape
        dxs = [torch.empty((M, w), dtype=torch.float32, device=W.device) if (need and w) else None
               for need, w in zip(ctx.need, ctx.widths)]
        dW = torch.empty_like(W)
        db = torch.empty(N, dtype=torch.float32, device=W.device)
        if ctx.mma:
            dev = W.device
            g_out = g_out.contiguous()
            hw, hs, hr = (ctypes.c_int64 * 3)(*ctx.widths), (ctypes.c_int64 * 3)(*ctx.strides), (ctypes.c_int64 * 3)(*ctx.rows)
            hd = (ctypes.c_int64 * 3)(*[0 if t is None else t.stride(0) for t in dxs])
            # the two products are independent: the weight gradient runs on an auxiliary stream beside the input gradient the rest of
            # the backward waits for
            cur = torch.cuda.current_stream(dev)
            aux = _aux_stream(dev) if _AUX_STREAM else None

            def weight_grad():
                _lib.call("igcn_catlin_mma_bwd_dw", _lib_ptr_strided(cs[0]), _lib_ptr_strided(cs[1]), _lib_ptr_strided(cs[2]),
                          ctypes.addressof(hw), ctypes.addressof(hs), ctypes.addressof(hr), _lib.ptr(out), _lib.ptr(g_out), M, N, K,
                          int(ctx.relu), _lib.ptr(dW), _lib.ptr(db), _lib.stream(), tag="cat_linear_bwd_w_mma[M=%d,N=%d,K=%d]" % (M, N, K),
                          nbytes=4 * (M * K + N * K + 2 * M * N))

            with torch.cuda.device(dev):
                if aux is not None:
                    aux.wait_stream(cur)
                    with torch.cuda.stream(aux):
                        weight_grad()
                    for t in [g_out, out, dW, db] + [t for t in cs if t is not None]:
                        t.record_stream(aux)
                if any(d is not None for d in dxs):
                    _lib.call("igcn_catlin_mma_bwd_dx", ctypes.addressof(hw), _lib.ptr(W), _lib.ptr(out), _lib.ptr(g_out), M, N, K,
                              int(ctx.relu), _lib.ptr(dxs[0]), _lib.ptr(dxs[1]), _lib.ptr(dxs[2]), ctypes.addressof(hd), _lib.stream(),
                              tag="cat_linear_bwd_x_mma[M=%d,N=%d,K=%d]" % (M, N, K),
                              nbytes=4 * (N * K + 2 * M * N + sum(M * w for w, d in zip(ctx.widths, dxs) if d is not None)))
                if aux is not None:
                    cur.wait_stream(aux)
                else:
                    weight_grad()
            for i, rp in enumerate(ctx.reps):                # a repeated source collects the gradient of every repetition
                if rp > 1 and dxs[i] is not None:
                    h = M // rp
                    acc = dxs[i][:h]
                    for k in range(1, rp):
                        acc = acc + dxs[i][k * h:(k + 1) * h]
                    dxs[i] = acc
            return dxs[0], dxs[1], dxs[2], dW, db, None
        if ctx.tc:
            dev = W.device
            g_out = g_out.contiguous()
            mask = out if ctx.relu else None
            gz = torch.empty((2, M, _pad4(N)), dtype=torch.float32, device=dev)          # gZ          (M, N): A of dX
            gzt = torch.empty((2, N, _pad4(M)), dtype=torch.float32, device=dev)         # gZ^T        (N, M): A of dW
            wt = torch.empty((2, K, _pad4(N)), dtype=torch.float32, device=dev)          # W^T         (K, N): B of dX
            xt = torch.empty((2, K + 1, _pad4(M)), dtype=torch.float32, device=dev)      # [X | 1]^T (K+1, M): B of dW (ones row -> d bias)
            # operands of the input gradient (small: gZ and W^T) and of the weight gradient (large: gZ^T and the transposed sources)
            jobs_x = [_job(g_out, gz, M, N, N, mask=mask), _job(W, wt, N, K, K, transpose=True)]
            jobs_w = [_job(g_out, gzt, M, N, N, transpose=True, mask=mask)]
            off = 0
            for t, w, ld, rp in zip(cs, ctx.widths, ctx.strides, ctx.reps):
                if t is not None and w:
                    for k in range(rp):
                        jobs_w.append(_job(t, xt, M // rp, w, ld, row_off=off, col_off=k * (M // rp), transpose=True))
                off += w
            jobs_w.append(_job(None, xt, M, 1, 1, row_off=K, transpose=True))

            def weight_grad():
                for i in range(0, len(jobs_w), 8):           # igcn_tc_split takes up to 8 jobs per launch
                    _tc_split(jobs_w[i:i + 8], dev)
                _tc_gemm(gzt, xt, N, K + 1, M, [dW, db], [K, 1], [K, 1], tag="cat_linear_bwd_w_tc")

            # the two products are independent: the weight gradient -- with its (much larger) operand preparation -- runs on an
            # auxiliary stream, so the input gradient the rest of the backward waits for is queued behind two small split jobs only
            cur = torch.cuda.current_stream(dev)
            aux = _aux_stream(dev) if _AUX_STREAM else None
            if aux is not None and aux == cur:       # this head's forward already ran on the auxiliary stream (forward_pair)
                aux = _aux_stream(dev, 1)
            if any(d is not None for d in dxs):          # queued first: the rest of the backward waits for it
                _tc_split(jobs_x, dev)
                _tc_gemm(gz, wt, M, K, N, dxs, ctx.widths, [0 if d is None else d.stride(0) for d in dxs], tag="cat_linear_bwd_x_tc")
            if aux is not None:
                # the weight-gradient side starts when the input-gradient GEMM has been issued, not beside it: its operand split fills
                # every SM with small CTAs, and the GEMM (one CTA per SM, all of its shared memory) then waits for the split to drain
                # -- measured on the device timeline as 8 us added to the path the attention backward waits for
                entry = torch.cuda.Event()
                entry.record(cur)
                aux.wait_event(entry)
                with torch.cuda.stream(aux):
                    weight_grad()
                for t in [g_out, gzt, xt, dW, db] + ([mask] if mask is not None else []) + [t for t in cs if t is not None]:
                    t.record_stream(aux)
                _join_or_defer(cur, aux, W)
            else:
                weight_grad()
            for i, rp in enumerate(ctx.reps):                # a repeated source collects the gradient of every repetition
                if rp > 1 and dxs[i] is not None:
                    h = M // rp
                    acc = dxs[i][:h]
                    for k in range(1, rp):
                        acc = acc + dxs[i][k * h:(k + 1) * h]
                    dxs[i] = acc
            return dxs[0], dxs[1], dxs[2], dW, db, None
        hw, hs = (ctypes.c_int64 * 3)(*ctx.widths), (ctypes.c_int64 * 3)(*ctx.strides)
        hd = (ctypes.c_int64 * 3)(*[0 if t is None else t.stride(0) for t in dxs])
        with torch.cuda.device(W.device):
            _lib.call("igcn_cat_linear_bwd", _lib_ptr_strided(cs[0]), _lib_ptr_strided(cs[1]), _lib_ptr_strided(cs[2]),
                      ctypes.addressof(hw), ctypes.addressof(hs), _lib.ptr(W), _lib.ptr(out), _lib.ptr(g_out.contiguous()), M, N, K,
                      int(ctx.relu), _lib.ptr(dxs[0]), _lib.ptr(dxs[1]), _lib.ptr(dxs[2]), ctypes.addressof(hd), _lib.ptr(dW), _lib.ptr(db),
                      _lib.stream(), tag="cat_linear_bwd[M=%d,N=%d,K=%d]" % (M, N, K),
                      nbytes=4 * (M * K + 2 * N * K + 2 * M * N + sum(M * w for w, d in zip(ctx.widths, dxs) if d is not None)))
        return dxs[0], dxs[1], dxs[2], dW, db, None


def _lib_ptr_strided(t):
    """device pointer of a 2-D tensor whose rows may be strided (last dim contiguous)."""
    return None if t is None else t.data_ptr()


def cat_linear(xs, weight, bias, relu=True):
    """relu(cat(xs, -1) @ weight.T + bias) without building the concatenation; xs: up to three (M, w_i) tensors."""
    xs = list(xs) + [None] * (3 - len(xs))
    return _CatLinearFn.apply(xs[0], xs[1], xs[2], weight, bias, relu)


class _CrossAttnFn(torch.autograd.Function):
    """relu: 0 = raw attention output, 1 = relu(attn), 2 = (q + relu(attn)) / 2 (the fusion average, tensor-core kernels only)."""

    @staticmethod
    def forward(ctx, q, kv, in_w, in_b, out_w, out_b, heads: int, relu: int):
        _lib.require_cuda(q, kv, in_w, in_b, out_w, out_b)
        c = lambda t: t.contiguous().float()
        q, kv, in_w, in_b, out_w, out_b = c(q), c(kv), c(in_w), c(in_b), c(out_w), c(out_b)
        B, R, E = q.shape
        M = kv.shape[1]
        out = torch.empty_like(q)
        lib = _lib.lib()
        # table-driven kernels (csrc/cross_attn_mma2.cuh) for the reference's shape: the per-graph key tables are saved for the backward
        tab = None
        if lib.igcn_cross_attn_v2_supported(R, M, E, heads):
            tab = torch.empty((B, lib.igcn_cross_attn_v2_tab_floats(M, heads)), dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            if tab is not None:
                _lib.call("igcn_cross_attn_v2_fwd", _lib.ptr(q), _lib.ptr(kv), _lib.ptr(in_w), _lib.ptr(in_b), _lib.ptr(out_w), _lib.ptr(out_b),
                          B, R, M, E, heads, int(relu), _lib.ptr(out), _lib.ptr(tab), _lib.stream(),
                          tag="cross_attn_fwd[R=%d,M=%d,E=%d]" % (R, M, E), nbytes=4 * (2 * B * R * E + B * M * E + 4 * E * E + 4 * E))
            else:
                _lib.call("igcn_cross_attn_fwd", _lib.ptr(q), _lib.ptr(kv), _lib.ptr(in_w), _lib.ptr(in_b), _lib.ptr(out_w), _lib.ptr(out_b),
                          B, R, M, E, heads, int(relu), _lib.ptr(out), _lib.stream(), tag="cross_attn_fwd[R=%d,M=%d,E=%d]" % (R, M, E),
                          nbytes=4 * (2 * B * R * E + B * M * E + 4 * E * E + 4 * E))
        ctx.heads, ctx.relu, ctx.v2 = heads, int(relu), tab is not None
        if tab is not None:
            ctx.save_for_backward(q, kv, in_w, in_b, out_w, out_b, out, tab)
        else:
            ctx.save_for_backward(q, kv, in_w, in_b, out_w, out_b, out)
        return out

    @staticmethod
    def backward(ctx, g):
        if ctx.v2:
            q, kv, in_w, in_b, out_w, out_b, out, tab = ctx.saved_tensors
        else:
            q, kv, in_w, in_b, out_w, out_b, out = ctx.saved_tensors
        lib = _lib.lib()
        B, R, E = q.shape
        M = kv.shape[1]
        P = lib.igcn_cross_attn_param_count(E)
        n_cta = lib.igcn_cross_attn_v2_bwd_ctas(B) if ctx.v2 else lib.igcn_cross_attn_bwd_ctas(B, R, M, E, ctx.heads)
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        partials = torch.empty((max(n_cta, 1), P), dtype=torch.float32, device=q.device)
        grads = torch.empty(P, dtype=torch.float32, device=q.device)
        if ctx.v2:
            work = torch.empty(max(lib.igcn_cross_attn_v2_work_floats(B, R, M, ctx.heads), 4), dtype=torch.float32, device=q.device)
            with torch.cuda.device(q.device):
                _lib.call("igcn_cross_attn_v2_bwd", _lib.ptr(q), _lib.ptr(kv), _lib.ptr(in_w), _lib.ptr(in_b), _lib.ptr(out_w), _lib.ptr(out_b),
                          _lib.ptr(out), _lib.ptr(g.contiguous()), _lib.ptr(tab), B, R, M, E, ctx.heads, int(ctx.relu), _lib.ptr(dq),
                          _lib.ptr(dkv), _lib.ptr(work), _lib.ptr(partials), n_cta, _lib.ptr(grads), _lib.stream(),
                          tag="cross_attn_bwd[R=%d,M=%d,E=%d]" % (R, M, E), nbytes=4 * (4 * B * R * E + 2 * B * M * E + 2 * (4 * E * E + 4 * E)))
            o1, o2, o3 = 3 * E * E, 3 * E * E + 3 * E, 4 * E * E + 3 * E
            return dq, dkv, grads[:o1].view(3 * E, E), grads[o1:o2], grads[o2:o3].view(E, E), grads[o3:], None, None
        with torch.cuda.device(q.device):
            _lib.call("igcn_cross_attn_bwd", _lib.ptr(q), _lib.ptr(kv), _lib.ptr(in_w), _lib.ptr(in_b), _lib.ptr(out_w), _lib.ptr(out_b),
                      _lib.ptr(out), _lib.ptr(g.contiguous()), B, R, M, E, ctx.heads, int(ctx.relu), _lib.ptr(dq), _lib.ptr(dkv),
                      _lib.ptr(partials), n_cta, _lib.ptr(grads), _lib.stream(), tag="cross_attn_bwd[R=%d,M=%d,E=%d]" % (R, M, E),
                      nbytes=4 * (4 * B * R * E + 2 * B * M * E + 2 * (4 * E * E + 4 * E)))
        o1, o2, o3 = 3 * E * E, 3 * E * E + 3 * E, 4 * E * E + 3 * E
        return dq, dkv, grads[:o1].view(3 * E, E), grads[o1:o2], grads[o2:o3].view(E, E), grads[o3:], None, None


def _check_mha(mha):
    if mha.in_proj_weight is None or mha.in_proj_bias is None or mha.bias_k is not None or mha.dropout != 0.0 or not mha.batch_first:
        raise RuntimeError("igcn_b200.cross_attention supports the reference configuration only "
                           "(packed in_proj with bias, no bias_kv, dropout 0, batch_first)")


def cross_attention(q, kv, mha: torch.nn.MultiheadAttention, relu=True):
    """relu(mha(q, kv, kv)[0]) for a batch_first nn.MultiheadAttention with packed in_proj (the reference's use)."""
    _check_mha(mha)
    return _CrossAttnFn.apply(q, kv, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, mha.num_heads,
                              1 if relu else 0)


def cross_attention_average(q, kv, mha: torch.nn.MultiheadAttention):
    """(q + relu(mha(q, kv, kv)[0])) / 2 -- the image/SNP fusion `out_z = (img_out + out_cross) / 2` of kernel/sgcn_img_snp.py
    folded into the attention epilogue (one kernel each way, the attention output itself is never materialised).  Falls back to
    the two-step form for shapes outside the tensor-core kernels."""
    _check_mha(mha)
    B, R, E = q.shape
    if q.is_cuda and _lib.lib().igcn_cross_attn_fused_average(R, kv.shape[1], E, mha.num_heads):
        return _CrossAttnFn.apply(q, kv, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, mha.num_heads, 2)
    return (q + cross_attention(q, kv, mha, relu=True)) * 0.5


class MaskBank(object):
    """All dropout masks of one forward pass from ONE kernel launch (igcn_dropout_masks).

    The first pass at a given batch size draws its masks with torch and records (name, shape, keep); every later pass
    fills one flat buffer with a single Philox launch and hands out views.  Masks are multiplicative scales (0 or 1/keep)."""

    def __init__(self, seed=None):
        self.plans = {}
        self.active = False
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self._rank_mixed = seed is not None          # an explicit seed is used as given (tests replay it)
        self.counter = None

    def _mix_rank(self):
        """Data-parallel ranks start from the same torch seed (identical replicas); their dropout masks must still differ."""
        if not self._rank_mixed:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.seed = (self.seed ^ ((dist.get_rank() + 1) * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF
            self._rank_mixed = True

    def begin_pass(self, key, device):
        import ctypes
        self.active, self.key, self.device = True, (key, str(device)), device
        self.recording, self.views = [], None
        plan = self.plans.get(self.key)
        if plan is None:
            return
        if self.counter is None or self.counter.device != device:
            self.counter = torch.zeros(1, dtype=torch.int64, device=device)
            self._mix_rank()
        names, shapes, ends, keeps = plan
        out = torch.empty(ends[-1], dtype=torch.float32, device=device)
        he, hk = (ctypes.c_int64 * len(ends))(*ends), (ctypes.c_float * len(keeps))(*keeps)
        with torch.cuda.device(device):
            _lib.call("igcn_dropout_masks", _lib.ptr(out), ctypes.addressof(he), ctypes.addressof(hk), len(ends), self.seed,
                      _lib.ptr(self.counter), _lib.stream())
        self.views, start = {}, 0
        for n, sh, e in zip(names, shapes, ends):
            self.views[n] = out[start:e].view(sh)
            start = e

    def get(self, name, shape, p):
        shape, keep = tuple(shape), 1.0 - p
        if self.views is not None:
            v = self.views.get(name)
            if v is not None and tuple(v.shape) == shape:
                return v
            self.plans.pop(self.key, None)           # shapes changed: re-record next pass
        m = torch.bernoulli(torch.full(shape, keep, device=self.device)) / keep
        self.recording.append((name, shape, keep))
        return m

    def end_pass(self):
        if self.active and self.views is None and self.recording and self.key not in self.plans:
            names = [r[0] for r in self.recording]
            if len(set(names)) == len(names) and len(names) <= 32:
                ends, tot = [], 0
                for _, sh, _ in self.recording:
                    n = 1
                    for d in sh:
                        n *= d
                    tot += n
                    ends.append(tot)
                self.plans[self.key] = (names, [r[1] for r in self.recording], ends, [r[2] for r in self.recording])
        self.active = False


class _GramFn(torch.autograd.Function):
    """G = S S^T (B x B) on the tensor cores (igcn_tc_gemm, 3xTF32); backward dS = (gG + gG^T) S is ONE product (the generic
    linear backward would spend a second, equally large one on the 'weight' side).  The dense contraction of
    OrthogonalConstraint (kernel/sgcn_img_snp.py:198-205) in its B x B form."""

    @staticmethod
    def forward(ctx, s):
        _lib.require_cuda(s)
        sc = s.contiguous().float()
        M, K = sc.shape
        out = torch.empty((M, M), dtype=torch.float32, device=s.device)
        if USE_TC:
            A = torch.empty((2, M, _pad4(K)), dtype=torch.float32, device=s.device)
            _tc_split([_job(sc, A, M, K, K)], s.device)
            _tc_gemm(A, A, M, M, K, [out], [M], [M], tag="gram_fwd_tc")
        else:
            import ctypes
            lib = _lib.lib()
            S = lib.igcn_cat_linear_splits(M, M, K)
            part = torch.empty((S, M, M), dtype=torch.float32, device=s.device)
            zero = torch.zeros(M, dtype=torch.float32, device=s.device)
            hw, hs = (ctypes.c_int64 * 3)(K, 0, 0), (ctypes.c_int64 * 3)(K, 0, 0)
            with torch.cuda.device(s.device):
                _lib.call("igcn_cat_linear_fwd", _lib.ptr(sc), None, None, ctypes.addressof(hw), ctypes.addressof(hs), _lib.ptr(sc), _lib.ptr(zero),
                          M, M, K, 0, _lib.ptr(part), S, _lib.ptr(out), _lib.stream(), tag="gram_fwd[B=%d,D=%d]" % (M, K),
                          nbytes=4 * (M * K + M * M))
        ctx.save_for_backward(sc, out)
        return out

    @staticmethod
    def backward(ctx, g):
        import ctypes
        sc, out = ctx.saved_tensors
        M, K = sc.shape
        gsym = (g + g.t()).contiguous()
        ds = torch.empty_like(sc)
        if USE_TC:
            # dS[i][c] = sum_j gsym[i][j] S^T[c][j]: A = gsym (M x M), B = S^T (K x M) straight out of the split pass
            ga = torch.empty((2, M, _pad4(M)), dtype=torch.float32, device=sc.device)
            st = torch.empty((2, K, _pad4(M)), dtype=torch.float32, device=sc.device)
            _tc_split([_job(gsym, ga, M, M, M), _job(sc, st, M, K, K, transpose=True)], sc.device)
            _tc_gemm(ga, st, M, K, M, [ds], [K], [K], tag="gram_bwd_tc")
            return ds
        hw, hs, hd = (ctypes.c_int64 * 3)(K, 0, 0), (ctypes.c_int64 * 3)(K, 0, 0), (ctypes.c_int64 * 3)(K, 0, 0)
        with torch.cuda.device(sc.device):
            _lib.call("igcn_cat_linear_bwd", _lib.ptr(sc), None, None, ctypes.addressof(hw), ctypes.addressof(hs), _lib.ptr(sc), _lib.ptr(out),
                      _lib.ptr(gsym), M, M, K, 0, _lib.ptr(ds), None, None, ctypes.addressof(hd), None, None, _lib.stream(),
                      tag="gram_bwd[B=%d,D=%d]" % (M, K), nbytes=4 * (2 * M * K + M * M))
        return ds


def gram(s):
    """s @ s.T for a (B, D) CUDA tensor."""
    return _GramFn.apply(s)


class _BnActFn(torch.autograd.Function):
    """mask * relu(BatchNorm1d(z)) in training mode, statistics per `groups` consecutive batch slices (igcn_bn_act_*)."""

    @staticmethod
    def forward(ctx, z, gamma, beta, mask, bn, groups: int, relu: bool):
        _lib.require_cuda(z, gamma, beta, mask)
        zc = z.contiguous().float()
        N, C = zc.shape[0], zc.shape[1]
        L = zc.numel() // (N * C)
        mc = None if mask is None else mask.expand_as(zc).contiguous().float()
        y = torch.empty_like(zc)
        stats = torch.empty((groups, C, 2), dtype=torch.float32, device=zc.device)
        rm, rv, nbt = bn.running_mean, bn.running_var, bn.num_batches_tracked
        mom = 0.1 if bn.momentum is None else float(bn.momentum)
        with torch.cuda.device(zc.device):
            _lib.call("igcn_bn_act_fwd", _lib.ptr(zc), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(mc), N, C, L, groups, float(bn.eps), mom,
                      int(relu), _lib.ptr(rm), _lib.ptr(rv), _lib.ptr(nbt), _lib.ptr(y), _lib.ptr(stats), _lib.stream(),
                      tag="bn_act_fwd[C=%d,L=%d]" % (C, L), nbytes=4 * zc.numel() * (2 + (mc is not None)))
        ctx.dims, ctx.groups, ctx.relu = (N, C, L), groups, bool(relu)
        ctx.save_for_backward(zc, gamma, beta, mc, stats)
        return y

    @staticmethod
    def backward(ctx, gy):
        zc, gamma, beta, mc, stats = ctx.saved_tensors
        N, C, L = ctx.dims
        dz = torch.empty_like(zc)
        dg = torch.empty(C, dtype=torch.float32, device=zc.device) if gamma is not None else None
        db = torch.empty(C, dtype=torch.float32, device=zc.device) if beta is not None else None
        with torch.cuda.device(zc.device):
            _lib.call("igcn_bn_act_bwd", _lib.ptr(zc), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(mc), _lib.ptr(stats), _lib.ptr(gy.contiguous()),
                      N, C, L, ctx.groups, int(ctx.relu), _lib.ptr(dz), _lib.ptr(dg), _lib.ptr(db), _lib.stream(),
                      tag="bn_act_bwd[C=%d,L=%d]" % (C, L), nbytes=4 * zc.numel() * (3 + (mc is not None)))
        return dz, dg, db, None, None, None, None


class _LinBnActFn(torch.autograd.Function):
    """mask * relu(BatchNorm1d_c(x W^T)) for x (N, C, K), bias-free W (L, K): the read-out Linear fused with the BatchNorm head that
    follows it (igcn_lin_bn_act_*; training mode, two stacked passes).  z = x W^T is never materialised."""

    @staticmethod
    def forward(ctx, x, W, gamma, beta, mask, bn, relu: bool):
        _lib.require_cuda(x, W, gamma, beta, mask)
        xc, Wc = x.contiguous().float(), W.contiguous().float()
        N, C, K = xc.shape
        L = Wc.shape[0]
        mc = None if mask is None else mask.reshape(N, C, -1).expand(N, C, L).contiguous().float()
        y = torch.empty((N, C, L), dtype=torch.float32, device=xc.device)
        stats = torch.empty((2, C, 2), dtype=torch.float32, device=xc.device)
        mom = 0.1 if bn.momentum is None else float(bn.momentum)
        with torch.cuda.device(xc.device):
            _lib.call("igcn_lin_bn_act_fwd", _lib.ptr(xc), _lib.ptr(Wc), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(mc), N, C, L, K, 2,
                      float(bn.eps), mom, int(relu), _lib.ptr(bn.running_mean), _lib.ptr(bn.running_var), _lib.ptr(bn.num_batches_tracked),
                      _lib.ptr(y), _lib.ptr(stats), _lib.stream(), tag="lin_bn_act_fwd[C=%d,K=%d,L=%d]" % (C, K, L),
                      nbytes=4 * (N * C * (K + L * (1 + (mc is not None)))))
        ctx.relu = bool(relu)
        ctx.save_for_backward(xc, Wc, gamma, beta, mc, stats)
        return y

    @staticmethod
    def backward(ctx, gy):
        xc, Wc, gamma, beta, mc, stats = ctx.saved_tensors
        N, C, K = xc.shape
        L = Wc.shape[0]
        dx = torch.empty_like(xc)
        dW = torch.empty_like(Wc)
        part = torch.empty((int(_lib.lib().igcn_lin_bn_act_partial_rows(N, C, L, K)), L * K), dtype=torch.float32, device=xc.device)
        dg = torch.empty(C, dtype=torch.float32, device=xc.device) if gamma is not None else None
        db = torch.empty(C, dtype=torch.float32, device=xc.device) if beta is not None else None
        with torch.cuda.device(xc.device):
            _lib.call("igcn_lin_bn_act_bwd", _lib.ptr(xc), _lib.ptr(Wc), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(mc), _lib.ptr(stats),
                      _lib.ptr(gy.contiguous().float()), N, C, L, K, 2, int(ctx.relu), _lib.ptr(dx), _lib.ptr(part), _lib.ptr(dW), _lib.ptr(dg),
                      _lib.ptr(db), _lib.stream(), tag="lin_bn_act_bwd[C=%d,K=%d,L=%d]" % (C, K, L),
                      nbytes=4 * (N * C * (2 * K + L * (1 + (mc is not None)))))
        return dx, dW, dg, db, None, None, None


def lin_bn_act(x, weight, bn: torch.nn.BatchNorm1d, mask=None, groups=1, relu=True):
    """mask * relu(bn(x @ weight.T)) for x (N, C, K) and a bias-free Linear weight (L, K), BatchNorm over the C axis: one fused launch
    each way when the shape fits (training mode, two stacked passes, K <= 8, L in {1, 32}); otherwise skinny_linear + bn_act.
    Returns (N, C, L)."""
    N, C, K = x.shape
    L = weight.shape[0]
    if (os.environ.get("IGCN_NO_LIN_BN", "") == "" and bn.training and bn.track_running_stats and x.is_cuda
            and _lib.lib().igcn_lin_bn_act_supported(N, C, L, K, groups)):
        return _LinBnActFn.apply(x, weight, bn.weight, bn.bias, mask, bn, relu)
    z = skinny_linear(x, weight) if (K <= 32 and L <= 64) else torch.nn.functional.linear(x, weight)
    if L == 1:                                            # BatchNorm1d over (N, C): the reference squeezes the last axis
        m2 = None if mask is None else mask.reshape(N, C)
        return bn_act(z.squeeze(-1), bn, m2, groups, relu).unsqueeze(-1)
    return bn_act(z, bn, None if mask is None else mask.reshape(N, C, -1).expand(N, C, L), groups, relu)


class _BnEvalActFn(torch.autograd.Function):
    """relu(BatchNorm1d(z)) with the running statistics (model.eval()): one elementwise launch each way (igcn_bn_eval_act)."""

    @staticmethod
    def forward(ctx, z, gamma, beta, rm, rv, eps: float, relu: bool):
        _lib.require_cuda(z, gamma, beta, rm, rv)
        zc = z.contiguous().float()
        N, C = zc.shape[0], zc.shape[1]
        L = zc.numel() // max(N * C, 1)
        y = torch.empty_like(zc)
        with torch.cuda.device(zc.device):
            _lib.call("igcn_bn_eval_act", _lib.ptr(zc), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(rm), _lib.ptr(rv), N, C, L, float(eps), int(relu),
                      None, _lib.ptr(y), _lib.stream(), tag="bn_eval_act[C=%d,L=%d]" % (C, L), nbytes=8 * zc.numel())
        ctx.dims, ctx.eps, ctx.relu = (N, C, L), float(eps), bool(relu)
        ctx.save_for_backward(zc, gamma, beta, rm, rv)
        return y

    @staticmethod
    def backward(ctx, gy):
        zc, gamma, beta, rm, rv = ctx.saved_tensors
        N, C, L = ctx.dims
        dz = torch.empty_like(zc)
        with torch.cuda.device(zc.device):
            _lib.call("igcn_bn_eval_act", _lib.ptr(zc), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(rm), _lib.ptr(rv), N, C, L, ctx.eps, int(ctx.relu),
                      _lib.ptr(gy.contiguous().float()), _lib.ptr(dz), _lib.stream(), tag="bn_eval_act_bwd", nbytes=12 * zc.numel())
        return dz, None, None, None, None, None, None          # eval mode: the affine parameters are constants


def bn_act(z, bn: torch.nn.BatchNorm1d, mask=None, groups=1, relu=True):
    """mask * relu(bn(z)) for a BatchNorm1d over (N,C) or (N,C,L) CUDA input, one launch each way.
    Training mode: batch statistics; groups > 1: z stacks that many passes along the batch, each gets its own statistics and
    running-buffer update, in order, exactly as `groups` successive module calls would (reference: kernel/go_model.py:117-146).
    Eval mode: the running statistics (no dropout mask): a per-channel affine map (the inference path)."""
    if not bn.track_running_stats:
        raise RuntimeError("igcn_b200.bn_act needs a BatchNorm1d that tracks running statistics")
    if not bn.training:
        if mask is not None:
            raise RuntimeError("igcn_b200.bn_act: a dropout mask in eval mode")
        return _BnEvalActFn.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, relu)
    return _BnActFn.apply(z, bn.weight, bn.bias, mask, bn, groups, relu)


class _MaskLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob, p_e, snps_prob, coef, eps):
        import ctypes
        _lib.require_cuda(prob, p_e, snps_prob)
        pc, ec, sc = prob.contiguous().float(), p_e.contiguous().float(), snps_prob.contiguous().float()
        lib = _lib.lib()
        nb = lib.igcn_reduce_blocks(max(pc.numel(), ec.numel(), sc.numel()))
        part = torch.empty(nb, dtype=torch.float32, device=pc.device)
        loss = torch.empty((), dtype=torch.float32, device=pc.device)
        hc = (ctypes.c_float * 4)(*coef)
        with torch.cuda.device(pc.device):
            _lib.call("igcn_mask_loss_fwd", _lib.ptr(pc), pc.numel(), _lib.ptr(ec) if ec.numel() else None, ec.numel(), _lib.ptr(sc), sc.numel(),
                      ctypes.addressof(hc), float(eps), _lib.ptr(part), nb, _lib.ptr(loss), _lib.stream(), tag="mask_loss_fwd",
                      nbytes=4 * (pc.numel() + ec.numel() + sc.numel()))
        ctx.coef, ctx.eps = tuple(coef), float(eps)
        ctx.save_for_backward(pc, ec, sc)
        return loss

    @staticmethod
    def backward(ctx, g):
        import ctypes
        pc, ec, sc = ctx.saved_tensors
        dp = torch.empty_like(pc) if ctx.needs_input_grad[0] else None
        de = torch.empty_like(ec) if ctx.needs_input_grad[1] else None
        ds = torch.empty_like(sc) if ctx.needs_input_grad[2] else None
        hc = (ctypes.c_float * 4)(*ctx.coef)
        with torch.cuda.device(pc.device):
            _lib.call("igcn_mask_loss_bwd", _lib.ptr(pc), pc.numel(), _lib.ptr(ec) if ec.numel() else None, ec.numel(), _lib.ptr(sc), sc.numel(),
                      ctypes.addressof(hc), ctx.eps, _lib.ptr(g.contiguous().float()), _lib.ptr(dp), _lib.ptr(de), _lib.ptr(ds), _lib.stream(),
                      tag="mask_loss_bwd", nbytes=8 * (pc.numel() + ec.numel() + sc.numel()))
        return dp, de, ds, None, None


def mask_loss(prob, p_e, snps_prob, hp, eps=1e-6):
    """loss_probability (kernel/sgcn_img_snp.py:153-181) of the raw node mask `prob`, the edge probabilities `p_e` and the raw
    SNP mask `snps_prob` as one fused reduction."""
    return _MaskLossFn.apply(prob, p_e, snps_prob, (hp.lamda_x_l1, hp.lamda_e_l1, hp.lamda_x_ent, hp.lamda_e_ent), eps)


class _LaplacianQuadFn(torch.autograd.Function):
    """scale * sum_h <S_h, (D - W) S_h> for a SYMMETRIC similarity W (B,B) with row sums d and `halves` row blocks S_h of s
    (halves*B, D).  csrc/laplacian.cu: because (D - W) 1 = 0 the columns of every block are centred first, s' = s - mean, and
    T = d .* s' - W s' -- the large diagonal term is exact fp32 and the tensor-core product W s' (ONE product for all blocks,
    N = halves * D) only carries a small correction, so the cancellation that cost ~1e-4 in round 1 is gone.  T serves the value
    (<s', T>) and the gradient (2 * scale * T; the centring needs no backward since (D - W) 1 = 0).  W = None: all ones
    (isSoftSimilarity=False): W s' = 0, d = B, no product at all."""

    @staticmethod
    def forward(ctx, s, W, d, scale: float, halves: int):
        _lib.require_cuda(s, W, d)
        sc = s.contiguous().float()
        MB, K = sc.shape
        M = MB // halves
        if M * halves != MB or not 1 <= halves <= 3 or (W is not None and (tuple(W.shape) != (M, M) or d.numel() != M)):
            raise RuntimeError("laplacian_quadratic: s is (%d, %d), W %s, halves=%d" % (MB, K, None if W is None else tuple(W.shape), halves))
        lib = _lib.lib()
        dev = sc.device
        t = torch.empty_like(sc)
        m = torch.empty((halves, K), dtype=torch.float32, device=dev)
        nb = lib.igcn_reduce_blocks(sc.numel())
        part = torch.empty(nb, dtype=torch.float32, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("igcn_col_mean", _lib.ptr(sc), M, K, halves, _lib.ptr(m), _lib.stream(), tag="col_mean", nbytes=4 * sc.numel())
        u = None
        if W is not None:
            Wc, dc = W.contiguous().float(), d.contiguous().float()
            u = torch.empty_like(sc)
            # U_h[i][c] = sum_j W[i][j] S'_h^T[c][j] on the tensor cores (3xTF32): A = W (M x M), B = [S'_0^T ; S'_1^T ...] (halves*K x M)
            la = torch.empty((2, M, _pad4(M)), dtype=torch.float32, device=dev)
            st = torch.empty((2, halves * K, _pad4(M)), dtype=torch.float32, device=dev)
            jobs = [_job(Wc, la, M, M, M)]
            for h in range(halves):
                jobs.append(_job(sc[h * M:], st, M, K, K, row_off=h * K, transpose=True, sub=m[h]))
            _tc_split(jobs, dev)
            _tc_gemm(la, st, M, halves * K, M, [u[h * M:] for h in range(halves)], [K] * halves, [K] * halves, tag="laplacian_product_tc")
        with torch.cuda.device(dev):
            _lib.call("igcn_laplacian_finish", _lib.ptr(sc), _lib.ptr(m), _lib.ptr(dc) if W is not None else None, _lib.ptr(u), M, K, halves,
                      float(M), float(scale), _lib.ptr(t), _lib.ptr(part), nb, _lib.ptr(out), _lib.stream(), tag="laplacian_finish",
                      nbytes=4 * sc.numel() * (3 if u is not None else 2))
        ctx.scale = float(scale)
        ctx.save_for_backward(t)
        return out

    @staticmethod
    def backward(ctx, g):
        (t,) = ctx.saved_tensors
        ds = torch.empty_like(t)
        with torch.cuda.device(t.device):
            _lib.call("igcn_scale_by_scalar", _lib.ptr(t), _lib.ptr(g.contiguous().float()), 2.0 * ctx.scale, t.numel(), _lib.ptr(ds),
                      _lib.stream(), tag="scale_by_scalar", nbytes=8 * t.numel())
        return ds, None, None, None, None


def laplacian_quadratic(s, W, d, scale=1.0, halves=1):
    """scale * sum_h tr(s_h^T (diag(d) - W) s_h) over the `halves` row blocks s_h (B, D) of a (halves*B, D) CUDA tensor, for a
    SYMMETRIC (B, B) similarity W with row sums d (constants); W = d = None means the all-ones similarity."""
    return _LaplacianQuadFn.apply(s, W, d, scale, halves)


def rbf_similarity(t, gamma):
    """(W, d): W[i][j] = exp(-gamma ||t_i - t_j||^2) for the rows of t (B, R) -- rbf_kernel_torch of util/image_cluster.py:15-31 as
    kernel/sgcn_img_snp.py:188 calls it -- and its row sums, in one kernel pair (no cdist / exp / sum launches)."""
    _lib.require_cuda(t)
    tc = t.detach().contiguous().float()
    B, R = tc.shape
    W = torch.empty((B, B), dtype=torch.float32, device=tc.device)
    d = torch.empty(B, dtype=torch.float32, device=tc.device)
    with torch.cuda.device(tc.device):
        _lib.call("igcn_rbf_similarity", _lib.ptr(tc), B, R, float(gamma), _lib.ptr(W), _lib.ptr(d), _lib.stream(), tag="rbf_similarity",
                  nbytes=4 * (B * R + B * B + B))
    return W, d


class _SkinnyLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W):
        _lib.require_cuda(x, W)
        xc, Wc = x.contiguous().float(), W.contiguous().float()
        Lout, Kin = Wc.shape
        rows = xc.numel() // Kin
        z = torch.empty(xc.shape[:-1] + (Lout,), dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            _lib.call("igcn_skinny_linear_fwd", _lib.ptr(xc), _lib.ptr(Wc), rows, Kin, Lout, _lib.ptr(z), _lib.stream(),
                      tag="skinny_linear_fwd[K=%d,L=%d]" % (Kin, Lout), nbytes=4 * (rows * (Kin + Lout)))
        ctx.save_for_backward(xc, Wc)
        return z

    @staticmethod
    def backward(ctx, gz):
        xc, Wc = ctx.saved_tensors
        Lout, Kin = Wc.shape
        rows = xc.numel() // Kin
        n_cta = _lib.lib().igcn_skinny_linear_bwd_ctas(rows)
        dx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        part = torch.empty((n_cta, Lout * Kin), dtype=torch.float32, device=xc.device)
        dW = torch.empty_like(Wc)
        with torch.cuda.device(xc.device):
            _lib.call("igcn_skinny_linear_bwd", _lib.ptr(xc), _lib.ptr(Wc), _lib.ptr(gz.contiguous().float()), rows, Kin, Lout, _lib.ptr(dx),
                      _lib.ptr(part), n_cta, _lib.ptr(dW), _lib.stream(), tag="skinny_linear_bwd[K=%d,L=%d]" % (Kin, Lout),
                      nbytes=4 * (rows * (2 * Kin + Lout)))
        return dx, dW


def skinny_linear(x, weight):
    """x @ weight.T for a bias-free Linear with in_features <= 32 and out_features <= 64 applied to the last dim of a CUDA tensor."""
    return _SkinnyLinearFn.apply(x, weight)


class _SnpMaskPairFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, snps, snps_prob):
        _lib.require_cuda(snps, snps_prob)
        sc, pc = snps.contiguous().float(), snps_prob.contiguous().float().view(-1)
        B, S = sc.shape
        out = torch.empty((2 * B, S), dtype=torch.float32, device=sc.device)
        with torch.cuda.device(sc.device):
            _lib.call("igcn_snp_mask_pair_fwd", _lib.ptr(sc), _lib.ptr(pc), B, S, _lib.ptr(out), _lib.stream(), tag="snp_mask_pair_fwd",
                      nbytes=12 * B * S)
        ctx.save_for_backward(sc, pc)
        ctx.pshape = snps_prob.shape
        return out

    @staticmethod
    def backward(ctx, g):
        sc, pc = ctx.saved_tensors
        B, S = sc.shape
        dp = torch.empty(S, dtype=torch.float32, device=sc.device)
        with torch.cuda.device(sc.device):
            _lib.call("igcn_snp_mask_pair_bwd", _lib.ptr(sc), _lib.ptr(pc), _lib.ptr(g.contiguous().float()), B, S, _lib.ptr(dp), _lib.stream(),
                      tag="snp_mask_pair_bwd", nbytes=8 * B * S)
        return None, dp.view(ctx.pshape)


def snp_mask_pair(snps, snps_prob):
    """cat([snps, snps * sigmoid(snps_prob)], 0): the SNP input of the stacked plain / explain passes (snps is data, no gradient)."""
    return _SnpMaskPairFn.apply(snps, snps_prob)


class _HeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h1, m1, h2, m2, W1, b1, W2, b2):
        _lib.require_cuda(h1, h2, W1, b1, W2, b2, m1, m2)
        c = lambda t: None if t is None else t.contiguous().float()
        h1, m1, h2, m2, W1, b1, W2, b2 = c(h1), c(m1), c(h2), c(m2), c(W1), c(b1), c(W2), c(b2)
        rows, K = h1.shape
        C1, C2 = W1.shape[0], W2.shape[0]
        logp = torch.empty((rows, C1), dtype=torch.float32, device=h1.device)
        reg = torch.empty((rows, C2), dtype=torch.float32, device=h1.device)
        with torch.cuda.device(h1.device):
            _lib.call("igcn_heads_fwd", _lib.ptr(h1), _lib.ptr(m1), _lib.ptr(h2), _lib.ptr(m2), _lib.ptr(W1), _lib.ptr(b1), _lib.ptr(W2),
                      _lib.ptr(b2), rows, K, C1, C2, _lib.ptr(logp), _lib.ptr(reg), _lib.stream(), tag="heads_fwd",
                      nbytes=4 * rows * (2 * K * (1 + (m1 is not None)) + C1 + C2))
        ctx.save_for_backward(h1, m1, h2, m2, W1, b1, W2, b2, logp)
        ctx.set_materialize_grads(False)       # an unused head (e.g. the classifier when its loss weight is 0) costs nothing
        return logp, reg

    @staticmethod
    def backward(ctx, g_logp, g_reg):
        h1, m1, h2, m2, W1, b1, W2, b2, logp = ctx.saved_tensors
        rows, K = h1.shape
        C1, C2 = W1.shape[0], W2.shape[0]
        n_cta = _lib.lib().igcn_heads_bwd_ctas(rows)
        P = C1 * K + C1 + C2 * K + C2
        if g_logp is None and g_reg is None:
            return (None,) * 8
        dh1 = torch.empty_like(h1) if g_logp is not None else None
        dh2 = torch.empty_like(h2) if g_reg is not None else None
        part = torch.empty((n_cta, P), dtype=torch.float32, device=h1.device)
        grads = torch.empty(P, dtype=torch.float32, device=h1.device)
        c = lambda t: None if t is None else t.contiguous().float()
        with torch.cuda.device(h1.device):
            _lib.call("igcn_heads_bwd", _lib.ptr(h1), _lib.ptr(m1), _lib.ptr(h2), _lib.ptr(m2), _lib.ptr(W1), _lib.ptr(b1), _lib.ptr(W2),
                      _lib.ptr(b2), _lib.ptr(logp), _lib.ptr(c(g_logp)), _lib.ptr(c(g_reg)), rows, K, C1, C2, _lib.ptr(dh1), _lib.ptr(dh2),
                      _lib.ptr(part), n_cta, _lib.ptr(grads), _lib.stream(), tag="heads_bwd", nbytes=4 * rows * (4 * K + C1 + C2))
        o1, o2, o3 = C1 * K, C1 * K + C1, C1 * K + C1 + C2 * K
        return (dh1, None, dh2, None, grads[:o1].view(C1, K) if g_logp is not None else None, grads[o1:o2] if g_logp is not None else None,
                grads[o2:o3].view(C2, K) if g_reg is not None else None, grads[o3:] if g_reg is not None else None)


def output_heads(h1, m1, h2, m2, lin2: torch.nn.Linear, lin2_regr: torch.nn.Linear):
    """(log_softmax(lin2(h1 * m1)), lin2_regr(h2 * m2)) in one launch (kernel/sgcn_img_snp.py:290-291,300-301); the masks are the
    dropout scales of F.dropout (None in eval mode)."""
    return _HeadsFn.apply(h1, m1, h2, m2, lin2.weight, lin2.bias, lin2_regr.weight, lin2_regr.bias)


class _StepLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, reg2, target, xhat2, snps, loss_prob, quad, c_reg, c_rec, c_prob, c_clu):
        _lib.require_cuda(reg2, target, xhat2, snps, loss_prob, quad)
        c = lambda t: None if t is None else t.contiguous().float()
        reg2, target, xhat2, snps, loss_prob, quad = c(reg2), c(target), c(xhat2), c(snps), c(loss_prob), c(quad)
        n_reg, n_rec = target.numel(), snps.numel()
        if reg2.numel() != 2 * n_reg or xhat2.numel() != 2 * n_rec:
            raise RuntimeError("step_loss: stacked tensors must hold exactly two passes")
        out = torch.empty((), dtype=torch.float32, device=reg2.device)
        with torch.cuda.device(reg2.device):
            _lib.call("igcn_step_loss_fwd", _lib.ptr(reg2), _lib.ptr(target), n_reg, _lib.ptr(xhat2), _lib.ptr(snps), n_rec, _lib.ptr(loss_prob),
                      _lib.ptr(quad), c_reg, c_rec, c_prob, c_clu, _lib.ptr(out), _lib.stream(), tag="step_loss_fwd",
                      nbytes=4 * (3 * n_reg + 3 * n_rec))
        ctx.coef = (c_reg, c_rec, c_prob, c_clu)
        ctx.has = (loss_prob is not None, quad is not None)
        ctx.save_for_backward(reg2, target, xhat2, snps)
        return out

    @staticmethod
    def backward(ctx, g):
        reg2, target, xhat2, snps = ctx.saved_tensors
        d_reg, d_xhat = torch.empty_like(reg2), torch.empty_like(xhat2)
        d_lp = torch.empty((), dtype=torch.float32, device=reg2.device) if ctx.has[0] else None
        d_q = torch.empty((), dtype=torch.float32, device=reg2.device) if ctx.has[1] else None
        with torch.cuda.device(reg2.device):
            _lib.call("igcn_step_loss_bwd", _lib.ptr(reg2), _lib.ptr(target), target.numel(), _lib.ptr(xhat2), _lib.ptr(snps), snps.numel(),
                      _lib.ptr(g.contiguous().float()), *ctx.coef, _lib.ptr(d_reg), _lib.ptr(d_xhat), _lib.ptr(d_lp), _lib.ptr(d_q),
                      _lib.stream(), tag="step_loss_bwd", nbytes=4 * (5 * target.numel() + 5 * snps.numel()))
        return d_reg, None, d_xhat, None, d_lp, d_q, None, None, None, None


def step_loss_pair(reg2, target, xhat2, snps, loss_prob, quad, c_reg, c_rec, c_prob, c_clu):
    """c_reg * mse(reg2 vs target, both passes) + c_rec * sum((xhat2 - snps)^2) + c_prob * loss_prob + c_clu * quad, one launch."""
    return _StepLossFn.apply(reg2, target, xhat2, snps, loss_prob, quad, float(c_reg), float(c_rec), float(c_prob), float(c_clu))
