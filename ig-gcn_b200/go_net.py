"""`Gene_ontology_network` on the fused GO kernels (reference: kernel/go_model.py).

Same constructor, forward signature, return tuple and parameter names as the reference class, so state_dicts
are interchangeable.  Differences are inside:

  * index preparation works on the sparse COO of A / A_g directly (the reference densifies G x G,
    go_model.py:52-56,70-72) and additionally builds CSR/CSC forms for the kernels;
  * every hierarchy layer is ONE kernel batched over subjects (the reference loops over subjects in Python,
    go_model.py:236-244);
  * `n_snps` is taken from A_g (the reference hard-codes 54 in the unused `classification` head).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .ops import MaskBank


def _csr_csc(rows: np.ndarray, cols: np.ndarray, n_rows: int, n_cols: int):
    """rows/cols: nnz list sorted row-major.  Returns the int32 arrays the kernels need."""
    nnz = rows.size
    rowptr = np.zeros(n_rows + 1, np.int32)
    np.cumsum(np.bincount(rows, minlength=n_rows), out=rowptr[1:])
    order = np.argsort(cols, kind="stable")
    colptr = np.zeros(n_cols + 1, np.int32)
    np.cumsum(np.bincount(cols, minlength=n_cols), out=colptr[1:])
    return dict(rowptr=rowptr, col=cols.astype(np.int32), row_of=rows.astype(np.int32), colptr=colptr,
                crow=rows[order].astype(np.int32), cpos=order.astype(np.int32), n_rows=n_rows, n_cols=n_cols, nnz=int(nnz))


def _compact_rank(rows: np.ndarray) -> np.ndarray:
    """store_ind of the reference (go_model.py:161-168): rank of each nnz's row among the non-empty rows."""
    if rows.size == 0:
        return rows.astype(np.int64)
    return np.cumsum(np.concatenate([[0], (rows[1:] != rows[:-1]).astype(np.int64)]))


class _GoSpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, vals, g):
        _lib.require_cuda(data, vals)
        lib = _lib.lib()
        data, vals = data.contiguous().float(), vals.contiguous().float()
        B, C = data.shape[0], vals.shape[0]
        out = torch.empty((B, g["n_rows"], C), dtype=torch.float32, device=data.device)
        with torch.cuda.device(data.device):
            _lib.call("igcn_go_spmm_fwd", _lib.ptr(data), _lib.ptr(g["rowptr"]), _lib.ptr(g["col"]), _lib.ptr(vals), B, g["n_cols"],
                                      g["n_rows"], g["nnz"], C, _lib.ptr(out), _lib.stream(),
                      tag="go_spmm_fwd[rows=%d,cols=%d,C=%d]" % (g["n_rows"], g["n_cols"], C),
                      nbytes=4 * (B * g["n_cols"] + B * g["n_rows"] * C + g["nnz"] * (2 + C)))
        ctx.g = g
        ctx.need_in = data.requires_grad
        ctx.save_for_backward(data, vals)
        return out

    @staticmethod
    def backward(ctx, g_out):
        data, vals = ctx.saved_tensors
        g, lib = ctx.g, _lib.lib()
        B, C = data.shape[0], vals.shape[0]
        g_out = g_out.contiguous()
        d_in = torch.empty_like(data) if ctx.need_in else None
        d_vals = torch.empty_like(vals)
        # large hierarchies: transpose once, then coalesced dot products per nnz (see include/igcn_b200.h)
        ws = torch.empty((g["n_rows"] * C + g["n_cols"]) * B, dtype=torch.float32, device=data.device) if g["nnz"] >= 4096 and B >= 8 else None
        with torch.cuda.device(data.device):
            _lib.call("igcn_go_spmm_bwd", _lib.ptr(g_out), _lib.ptr(data), _lib.ptr(g["row_of"]), _lib.ptr(g["col"]), _lib.ptr(g["colptr"]),
                                      _lib.ptr(g["crow"]), _lib.ptr(g["cpos"]), _lib.ptr(vals), B, g["n_cols"], g["n_rows"], g["nnz"],
                                      C, _lib.ptr(d_in), _lib.ptr(d_vals), _lib.ptr(ws), _lib.stream(),
                      tag="go_spmm_bwd[rows=%d,cols=%d,C=%d]" % (g["n_rows"], g["n_cols"], C),
                      nbytes=4 * (2 * B * g["n_cols"] + B * g["n_rows"] * C + g["nnz"] * (4 + 2 * C)))
        return d_in, d_vals, None


class _GoLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Wa, Ws, u, v, gamma, beta, mask, g, attn, self_off, keep_from):
        _lib.require_cuda(x, Wa, Ws, gamma, beta)
        lib = _lib.lib()
        x = x.contiguous().float()
        B, m_in, din = x.shape
        dout = Wa.shape[0]
        m_row = g["n_rows"]
        Wa, Ws = Wa.contiguous().float(), Ws.contiguous().float()
        u = u.contiguous().float().view(-1) if attn else None
        v = v.contiguous().float().view(-1) if attn else None
        gamma, beta = gamma.contiguous().float(), beta.contiguous().float()
        mask = mask.contiguous().float().view(B, m_row) if mask is not None else None
        y = torch.empty((B, m_row - keep_from, dout), dtype=torch.float32, device=x.device)
        stats = torch.empty((B, 2 * dout), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("igcn_go_layer_fwd", _lib.ptr(x), _lib.ptr(Wa), _lib.ptr(Ws), _lib.ptr(u), _lib.ptr(v), _lib.ptr(gamma), _lib.ptr(beta),
                                       _lib.ptr(mask), _lib.ptr(g["rowptr"]), _lib.ptr(g["col"]), _lib.ptr(g["colptr"]), _lib.ptr(g["crow"]),
                                       _lib.ptr(g["cpos"]), B, m_in, m_row, g["nnz"], din, dout, int(attn), self_off, keep_from,
                                       _lib.ptr(y), _lib.ptr(stats), _lib.stream(), tag="go_layer_fwd[%s,M=%d]" % ("attn" if attn else "dec", m_row),
                      nbytes=4 * (B * m_in * din + B * (m_row - keep_from) * dout + (B * m_row if mask is not None else 0) + 2 * B * dout))
        ctx.g, ctx.attn, ctx.self_off, ctx.keep_from = g, attn, self_off, keep_from
        ctx.save_for_backward(x, Wa, Ws, u, v, gamma, beta, mask, stats)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, Wa, Ws, u, v, gamma, beta, mask, stats = ctx.saved_tensors
        g, attn, lib = ctx.g, ctx.attn, _lib.lib()
        B, m_in, din = x.shape
        dout, m_row = Wa.shape[0], g["n_rows"]
        P = lib.igcn_go_layer_param_count(din, dout, m_row)
        n_cta = lib.igcn_go_layer_bwd_ctas(B, din, dout, m_in, m_row, g["nnz"], int(attn))
        dx = torch.empty_like(x)
        partials = torch.empty((max(n_cta, 1), P), dtype=torch.float32, device=x.device)
        grads = torch.empty(P, dtype=torch.float32, device=x.device)
        gy = gy.contiguous()
        with torch.cuda.device(x.device):
            _lib.call("igcn_go_layer_bwd", _lib.ptr(x), _lib.ptr(Wa), _lib.ptr(Ws), _lib.ptr(u), _lib.ptr(v), _lib.ptr(gamma), _lib.ptr(beta),
                                       _lib.ptr(mask), _lib.ptr(g["rowptr"]), _lib.ptr(g["col"]), _lib.ptr(g["colptr"]), _lib.ptr(g["crow"]),
                                       _lib.ptr(g["cpos"]), B, m_in, m_row, g["nnz"], din, dout, int(attn), ctx.self_off, ctx.keep_from,
                                       _lib.ptr(stats), _lib.ptr(gy), _lib.ptr(dx), _lib.ptr(partials), n_cta, _lib.ptr(grads), _lib.stream(),
                      tag="go_layer_bwd[%s,M=%d]" % ("attn" if attn else "dec", m_row),
                      nbytes=4 * (2 * B * m_in * din + B * (m_row - ctx.keep_from) * dout + (B * m_row if mask is not None else 0) + 2 * B * dout))
        n = dout * din
        dWa, dWs = grads[:n].view(dout, din), grads[n:2 * n].view(dout, din)
        du = grads[2 * n:2 * n + 2 * dout].view(1, 2 * dout) if attn else None
        dv = grads[2 * n + 2 * dout:2 * n + 3 * dout].view(1, dout) if attn else None
        o = 2 * n + 3 * dout
        return dx, dWa, dWs, du, dv, grads[o:o + m_row], grads[o + m_row:o + 2 * m_row], None, None, None, None, None


class Gene_ontology_network(nn.Module):
    """Drop-in for kernel/go_model.py::Gene_ontology_network (ctor :24, forward :205)."""

    def __init__(self, A_g, A, in_f_dim, n_l, f_dim, pool_dim, l_dim, device, dim_snps_atten=5):
        super().__init__()
        self.device = device
        pool = [int(p) for p in pool_dim[0]]
        self.pool, self.n_l = pool, int(n_l)
        A = A.coalesce()
        A_g = A_g.coalesce()
        G, S = int(A_g.shape[0]), int(A_g.shape[1])
        self.G, self.S = G, S
        ar, ac = A.indices().cpu().numpy()          # A[parent, child] (row aggregates from col)
        keep = A.values().cpu().numpy() != 0
        ar, ac = ar[keep], ac[keep]
        # ---- encoder sub-adjacencies A[off_j:, off_j:]  (go_model.py:51-61) ---------------------------------
        self._graphs = {}
        self.n_loc_in, self.store_in = [], []
        off = 0
        for j in range(self.n_l):
            off = sum(pool[:j])
            m = (ar >= off) & (ac >= off)
            r, c = ar[m] - off, ac[m] - off
            self._graphs["enc%d" % j] = _csr_csc(r, c, G - off, G - off)
            self.n_loc_in.append(torch.from_numpy(np.stack([r, c]).astype(np.int64)))
            self.store_in.append(torch.from_numpy(_compact_rank(r)))
        # ---- decoder sub-adjacencies A^T[r0:, c0:]  (go_model.py:65-74) --------------------------------------
        tr, tc = ac, ar                              # A^T nnz, re-sorted row-major
        o = np.lexsort((tc, tr))
        tr, tc = tr[o], tc[o]
        self.n_loc_out, self.store_out = [], []
        for j in range(self.n_l):
            r0, c0 = sum(pool[: self.n_l - j - 1]), sum(pool[: self.n_l - j])
            m = (tr >= r0) & (tc >= c0)
            r, c = tr[m] - r0, tc[m] - c0
            self._graphs["dec%d" % j] = _csr_csc(r, c, G - r0, G - c0)
            self.n_loc_out.append(torch.from_numpy(np.stack([r, c]).astype(np.int64)))
            self.store_out.append(torch.from_numpy(_compact_rank(r)))
        # ---- SNP <-> GO incidence (go_model.py:78-88) ---------------------------------------------------------
        gr, gc = A_g.indices().cpu().numpy()
        self.i = torch.from_numpy(np.stack([gr, gc]).astype(np.int64))
        self.size = A_g.size()
        self._graphs["ag"] = _csr_csc(gr, gc, G, S)
        o = np.lexsort((gr, gc))
        self.i_D = torch.from_numpy(np.stack([gc[o], gr[o]]).astype(np.int64))
        self.size_D = torch.Size([S, G])
        self._graphs["ag_t"] = _csr_csc(gc[o], gr[o], S, G)
        nnz_g = int(gr.size)
        self.t = nn.ParameterList([nn.Parameter(torch.empty(nnz_g).normal_(1.0, 0.1)) for _ in range(in_f_dim)])
        self.t_D = nn.ParameterList([nn.Parameter(torch.empty(nnz_g).normal_(1.0, 0.1))])
        # ---- layers (names/shapes as go_model.py:91-157) -----------------------------------------------------
        f_dim = [in_f_dim] + list(f_dim)
        self.f_dim = f_dim
        n_l = self.n_l
        top = sum(pool) - sum(pool[0:n_l])
        self.w_inc = nn.ModuleList([nn.Linear(f_dim[i], f_dim[i + 1], bias=False) for i in range(n_l)])
        self.w_s_loop = nn.ModuleList([nn.Linear(f_dim[i], f_dim[i + 1], bias=False) for i in range(n_l)])
        self.w_att_s = nn.ModuleList([nn.Linear(f_dim[i + 1], 1, bias=False) for i in range(n_l)])
        self.w_att_s_act = nn.ModuleList([nn.Sigmoid() for _ in range(n_l)])
        self.G_B = nn.ModuleList([nn.LayerNorm(sum(pool[i:])) for i in range(n_l)])
        self.w_act = nn.ModuleList([nn.ReLU() for _ in range(n_l)])
        self.gcn_D = nn.ModuleList([nn.Dropout2d(0.4) for _ in range(n_l)])
        self.w_att_in = nn.ModuleList([nn.Linear(2 * f_dim[i + 1], 1, bias=False) for i in range(n_l)])
        self.w_att_in_act = nn.ModuleList([nn.Tanh() for _ in range(n_l)])
        self.w_out = nn.ModuleList([nn.Linear(f_dim[i], f_dim[i - 1], bias=False) for i in range(n_l, 0, -1)])
        self.w_s_loop_out = nn.ModuleList([nn.Linear(f_dim[i], f_dim[i - 1], bias=False) for i in range(n_l, 0, -1)])
        self.G_B_D = nn.ModuleList([nn.LayerNorm(sum(pool[i:])) for i in range(n_l - 1, -1, -1)])
        self.w_act_out = nn.ModuleList([nn.ReLU() for _ in range(n_l)])
        self.gcn_D_D = nn.ModuleList([nn.Dropout2d(0.4) for _ in range(n_l)])
        self.conc_for_attention = nn.Sequential(nn.Linear(f_dim[-1], dim_snps_atten, bias=False), nn.BatchNorm1d(top), nn.ReLU())
        self.conc = nn.Linear(f_dim[-1], 1, bias=False)
        self.B = nn.Sequential(nn.BatchNorm1d(top), nn.ReLU(), nn.Dropout(0.5))
        self.conc_D = nn.Linear(f_dim[0], 1, bias=False)
        self.B_D = nn.Sequential(nn.BatchNorm1d(sum(pool)), nn.ReLU(), nn.Dropout(0.5))
        self.latent = nn.Sequential(nn.Linear(top, 32, bias=False), nn.BatchNorm1d(32), nn.ReLU(), nn.Dropout(0.5),
                                    nn.Linear(32, l_dim, bias=False), nn.BatchNorm1d(l_dim), nn.ReLU())
        self.classification = nn.Sequential(nn.BatchNorm1d(l_dim + S), nn.ReLU(), nn.Dropout(0.5), nn.Linear(l_dim + S, 16, bias=False),
                                            nn.ReLU(), nn.Dropout(0.3), nn.Linear(16, 1, bias=True), nn.Sigmoid())
        self._dev_graphs = {}
        self._groups = 1
        self.dropout_masks = None     # test hook: dict name -> multiplicative scale tensor (oracle.GO_MASK_NAMES)
        self.mask_bank = MaskBank()   # all masks of a pass from one kernel launch (shared with the enclosing model)
        self.atten_ready = None       # optional torch.cuda.Event recorded as soon as atten_out is enqueued
        self.branch_stream = None     # optional torch.cuda.Stream for the decoder branch (runs beside the latent read-outs)
        self.latent_stream = None     # optional torch.cuda.Stream for the latent read-outs (run beside the attention tokens' read-out)
        self.decoder_joined = True    # False after a forward that left x_D on branch_stream: the caller joins that stream

    # graph index tensors follow the module's device lazily (they are not parameters / state_dict entries)
    def _g(self, name, dev):
        key = (name, str(dev))
        g = self._dev_graphs.get(key)
        if g is None:
            h = self._graphs[name]
            g = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in h.items()}
            self._dev_graphs[key] = g
        return g

    def _mask(self, name, shape, p, dev):
        if not self.training:
            return None
        if self.dropout_masks is not None:
            m = self.dropout_masks[name].to(dev).float()
            return m.expand(shape) if m.numel() == 1 else m.reshape(shape)
        return self.mask_bank.get(name, shape, p)

    def _drop(self, name, t, p):
        m = self._mask(name, t.shape, p, t.device)
        return t if m is None else t * m

    def _bn_act(self, bn, z, name=None, p=0.0):
        """dropout(relu(bn(z))) as one fused launch (ops.bn_act): batch statistics + mask in training mode, the running statistics
        in eval mode (the inference path of eval_acc / eval_loss / eval_scores)."""
        mask = self._mask(name, z.shape, p, z.device) if (name and self.training) else None
        return ops.bn_act(z, bn, mask, self._groups, relu=True)

    def _lin_bn(self, lin, bn, x, name=None, p=0.0):
        """dropout(relu(bn(lin(x)))) of a per-node read-out (x (N, C, K) -> (N, C, L)) as ONE fused launch each way (ops.lin_bn_act)."""
        N, C = x.shape[0], x.shape[1]
        L = lin.weight.shape[0]
        mshape = (N, C) if L == 1 else (N, C, L)
        mask = self._mask(name, mshape, p, x.device) if (name and self.training) else None
        return ops.lin_bn_act(x, lin.weight, bn, mask, self._groups, relu=True)

    @staticmethod
    def _lin(mod, x):
        """Bias-free read-out projection; the skinny shapes (in <= 32, out <= 64) run on their own kernels (glue.cu); anything larger
        is a plain library GEMM (nn.Linear)."""
        w = mod.weight
        if mod.bias is None and w.shape[1] <= 32 and w.shape[0] <= 64:
            return ops.skinny_linear(x, w)
        return mod(x)

    def forward(self, data, T=None, device=None, groups=1):
        """groups=2: `data` holds the plain pass and the explain pass stacked along the batch (train.step_loss)."""
        dev = data.device
        n_l, pool = self.n_l, self.pool
        self._groups = groups
        own_pass = self.training and self.dropout_masks is None and not self.mask_bank.active
        if own_pass:
            self.mask_bank.begin_pass(data.shape[0], dev)
        # SNP -> GO encode: (B,S) -> (B,G,in_f_dim)
        x = _GoSpmmFn.apply(data, torch.stack(list(self.t)), self._g("ag", dev))
        for j in range(n_l):
            g = self._g("enc%d" % j, dev)
            mask = self._mask("go_enc%d" % j, (x.shape[0], g["n_rows"]), 0.4, dev)
            x = _GoLayerFn.apply(x, self.w_inc[j].weight, self.w_s_loop[j].weight, self.w_att_in[j].weight, self.w_att_s[j].weight,
                                 self.G_B[j].weight, self.G_B[j].bias, mask, g, True, 0, pool[j])
        ls = self.latent_stream if (self.branch_stream is not None and x.is_cuda) else None
        if ls is not None:                        # the fusion heads wait for `latent`: its six kernels start right here, on their own stream
            ls.wait_stream(torch.cuda.current_stream(dev))
        # three consumers of the encoder output (attention tokens, decoder, latent read-outs): their gradients are summed once
        x_att, x_dec, x_lat = ops.fan_out(x, 3)
        atten_out = self._lin_bn(self.conc_for_attention[0], self.conc_for_attention[1], x_att)
        if self.atten_ready is not None:          # lets a caller on another stream start the cross attention before the decoder is done
            self.atten_ready.record(torch.cuda.current_stream(dev))
        def decoder(x):
            for j in range(n_l):
                g = self._g("dec%d" % j, dev)
                mask = self._mask("go_dec%d" % j, (x.shape[0], g["n_rows"]), 0.4, dev)
                x = _GoLayerFn.apply(x, self.w_out[j].weight, self.w_s_loop_out[j].weight, None, None, self.G_B_D[j].weight,
                                     self.G_B_D[j].bias, mask, g, False, pool[n_l - j - 1], 0)
            out_D = self._lin_bn(self.conc_D, self.B_D[0], x, "go_BD", 0.5).squeeze(-1)
            return _GoSpmmFn.apply(out_D, self.t_D[0].unsqueeze(0), self._g("ag_t", dev)).squeeze(-1)

        def latent_head(x):
            inp_out = self._lin_bn(self.conc, self.B[0], x, "go_B", 0.5).squeeze(-1)
            h = self._bn_act(self.latent[1], self._lin(self.latent[0], inp_out), "go_latent", 0.5)
            return self._bn_act(self.latent[5], self._lin(self.latent[4], h))

        # The decoder (-> x_D, needed by the reconstruction loss only) and the latent read-outs (-> the fusion heads) both hang off
        # the encoder output and are ~6 kernels each: on one stream the heads waited for the decoder.  With a branch stream they
        # run side by side (and, because autograd replays a node on the stream of its forward, so do their backward chains).
        br = self.branch_stream
        if br is not None and x.is_cuda:
            cur = torch.cuda.current_stream(dev)
            br.wait_stream(cur)
            with torch.cuda.stream(br):
                x_D = decoder(x_dec)
            x.record_stream(br)
            if ls is not None:
                with torch.cuda.stream(ls):
                    latent = latent_head(x_lat)
                x.record_stream(ls)
            else:
                latent = latent_head(x_lat)
            self.decoder_joined = False
        else:
            x_D = decoder(x_dec)
            latent = latent_head(x_lat)
            self.decoder_joined = True
        if own_pass:
            self.mask_bank.end_pass()
        return latent, x_D, [torch.zeros(3, device=dev)], atten_out
