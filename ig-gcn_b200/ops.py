"""autograd wrappers around the C-ABI kernels (include/igcn_b200.h).  CUDA tensors only."""
from __future__ import annotations

import torch

from . import _lib
from .data import GraphCSR


def pack_layer_params(weights, biases):
    """[W_1 (H,F0) | b_1 | W_2 (H,H) | b_2 | ...] -- the `wb` layout of igcn_sgcn_encoder_*."""
    parts = []
    for w, b in zip(weights, biases):
        parts.append(w.reshape(-1))
        parts.append(b.reshape(-1))
    return torch.cat(parts) if parts else None


class _SGCNEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, prob, prob_bias, wb, csr: GraphCSR, L: int, H: int, want_pe: bool, relu: bool = True):
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
        out = torch.empty((B, R, L * H), dtype=torch.float32, device=x.device)
        p_e = torch.empty(csr.E, dtype=torch.float32, device=x.device) if (explain and want_pe) else None
        with torch.cuda.device(x.device):
            _lib.call("igcn_sgcn_encoder_fwd", _lib.ptr(x), _lib.ptr(csr.rowptr_t), _lib.ptr(csr.csr_src), _lib.ptr(csr.csr_w),
                                           _lib.ptr(probc), _lib.ptr(pbc), _lib.ptr(wbc), B, R, F0, H, L, csr.max_eg, int(relu),
                                           _lib.ptr(out), _lib.ptr(p_e), _lib.stream(),
                      tag="sgcn_encoder_fwd[%s,L=%d]" % ("explain" if explain else "plain", L))
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
                      tag="sgcn_encoder_bwd[%s,L=%d]" % ("explain" if ctx.explain else "plain", L))
        nwb = P - R * F0 - 2 * F0
        d_wb = grads[:nwb] if wb is not None else None
        d_prob = grads[nwb:nwb + R * F0].view(R, F0) if ctx.explain else None
        d_pb = grads[nwb + R * F0:].view(2 * F0, 1) if ctx.explain else None
        return dx, d_prob, d_pb, d_wb, None, None, None, None, None


def sgcn_encoder(x, csr: GraphCSR, weights, biases, prob=None, prob_bias=None, want_pe=False, relu=True):
    """Fused SGCN encoder. Returns (out (B,R,L*H), p_e (E,) in CSR-slot order or empty)."""
    L = len(weights)
    H = weights[0].shape[0] if L else 0
    wb = pack_layer_params(weights, biases)
    return _SGCNEncoderFn.apply(x, prob, prob_bias, wb, csr, L, H, want_pe, relu)


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
                      _lib.ptr(out), _lib.stream())
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
                      _lib.ptr(partials), n_cta, _lib.ptr(grads), _lib.stream())
        o = H * Fin
        sh = ctx.shapes
        return (dx, d_ea, grads[:o].view(H, Fin), grads[o:o + H].view(sh[0]), grads[o + H:o + 2 * H].view(sh[1]),
                grads[o + 2 * H:o + 3 * H].view(sh[2]), grads[o + 3 * H:o + 4 * H].view(sh[3]), grads[o + 4 * H:o + 5 * H].view(sh[4]),
                None, None)


def gat_conv(x, csr: GraphCSR, edge_attr_csr, W, att_src, att_dst, lin_edge, att_edge, bias, negative_slope=0.2):
    """Fused GATConv(heads=1, edge_dim=1). edge_attr_csr is per CSR slot (edge_attr[csr.csr_perm])."""
    return _GATConvFn.apply(x, edge_attr_csr, W, att_src, att_dst, lin_edge, att_edge, bias, csr, negative_slope)
