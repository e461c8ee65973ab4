"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (plain torch on the host, any float dtype; autograd supplies the
gradient oracle) of the IG-GCN graph-convolution hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg
may import this file.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the pin is made here: `tests/golden/make_golden.py`
runs the reference's own UNMODIFIED files (kernel/sgcn_img_snp.py, kernel/sgcn.py,
kernel/go_model.py, batch.py, dataloader.py) on CPU -- on top of `oracle/shim`,
the restatement of the un-vendored torch_geometric 2.0.2 / torch_scatter 2.0.9 --
and commits inputs, parameters, outputs and gradients under tests/golden/.
`tests/test_oracle_golden.py` checks every function below against those files.
The shim itself restates third-party code that is absent from /root/reference
(environment.yml:183,211); its semantics are anchored on SURVEY.md section 3.4.

All functions are written in functional form over a flat parameter dict `P`
that uses the reference's state_dict names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# a1. collation  (batch.py:24-123 via dataloader.py:26-29) -- integer work, bit exact
# --------------------------------------------------------------------------------------
def collate(sub: dict, idx) -> Dict[str, np.ndarray]:
    """Concatenate the graphs `idx` of a packed subject set into one disconnected graph.
    Node ids of graph i are shifted by the cumulative node count (batch.py:54-55, Data.__inc__
    for keys containing 'index'), `batch` is full((R,), i) per graph (batch.py:96-99),
    1-row tensors stack along dim 0 and 1-D tensors concatenate (batch.py:104-108)."""
    idx = np.asarray(idx, dtype=np.int64)
    R = sub["x"].shape[1]
    ptr = sub["edge_ptr"]
    src, dst, w = [], [], []
    for k, g in enumerate(idx):
        e0, e1 = int(ptr[g]), int(ptr[g + 1])
        src.append(sub["edge_src"][e0:e1] + k * R)
        dst.append(sub["edge_dst"][e0:e1] + k * R)
        w.append(sub["edge_attr"][e0:e1])
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if len(xs) else np.zeros((0,), dt)
    return dict(
        x=sub["x"][idx].reshape(-1, sub["x"].shape[2]).astype(np.float32),
        edge_index=np.stack([cat(src, np.int64), cat(dst, np.int64)]),
        edge_attr=cat(w, np.float32),
        batch=np.repeat(np.arange(idx.size, dtype=np.int64), R),
        snps_feat=sub["snps_feat"][idx].astype(np.float32),
        y=sub["y"][idx].astype(np.int64),
        clini_score=sub["clini_score"][idx].reshape(-1).astype(np.float32),
        tsne_fdim=sub["tsne_fdim"][idx].astype(np.float32),
        clust_y=sub["clust_y"][idx].astype(np.int64),
        sbjID=sub["sbjID"][idx].astype(np.int64),
    )


def target_sorted_csr(edge_index: np.ndarray, num_nodes: int):
    """The device layout contract: in-edges grouped by TARGET, ties in original edge order
    (stable), i32.  Returns (rowptr (N+1,), src (E,), perm (E,)) with perm[k] = original edge id."""
    dst = edge_index[1]
    perm = np.argsort(dst, kind="stable").astype(np.int32)
    rowptr = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(np.bincount(dst, minlength=num_nodes), out=rowptr[1:])
    return rowptr, edge_index[0][perm].astype(np.int32), perm


def source_sorted_csr(edge_index: np.ndarray, num_nodes: int):
    """Out-edges grouped by SOURCE (the transposed operator used by the backward pass).
    Returns (colptr (N+1,), pos (E,)) where pos[k] is the position, in the target-sorted
    list, of the k-th out-edge."""
    src, dst = edge_index[0], edge_index[1]
    perm_t = np.argsort(dst, kind="stable")
    inv = np.empty_like(perm_t)
    inv[perm_t] = np.arange(perm_t.size)
    perm_s = np.argsort(src, kind="stable")
    colptr = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(np.bincount(src, minlength=num_nodes), out=colptr[1:])
    return colptr, inv[perm_s].astype(np.int32)


# --------------------------------------------------------------------------------------
# a2. importance masks  (kernel/sgcn_img_snp.py:133-151 ; kernel/sgcn.py:74-85)
# --------------------------------------------------------------------------------------
def cal_probability(P, x: Tensor, edge_index: Tensor, edge_weight: Tensor, rois: int,
                    snps: Optional[Tensor] = None, prefix: str = ""):
    prob, pb = P[prefix + "prob"], P[prefix + "prob_bias"]
    f0 = x.shape[1]
    xm = (x.view(-1, rois, f0) * prob).reshape(-1, f0)            # RAW prob, no sigmoid (:136-137)
    z = xm[edge_index[0]] @ pb[:f0, 0] + xm[edge_index[1]] @ pb[f0:, 0]   # source half first (:141)
    pe = torch.sigmoid(z)
    out = dict(x=xm, w=edge_weight * pe, p_e=pe)
    if snps is not None:
        sp = torch.sigmoid(P[prefix + "snps_prob"])
        out.update(snps=snps * sp, snps_p=sp)
    return out


# --------------------------------------------------------------------------------------
# a3. GCNConv  (PyG 2.0.2 semantics)
# --------------------------------------------------------------------------------------
def gcn_norm(edge_index: Tensor, w: Tensor, n: int):
    """Self-loop merge (existing loop keeps its weight, else 1), in-degree by target,
    D^-1/2 A D^-1/2.  Returns the re-built edge list (non-loops in order, loops last)."""
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    loop_w = torch.ones(n, dtype=w.dtype)
    loop_w = loop_w.index_put((src[~keep],), w[~keep])
    s2 = torch.cat([src[keep], torch.arange(n)])
    d2 = torch.cat([dst[keep], torch.arange(n)])
    w2 = torch.cat([w[keep], loop_w])
    deg = torch.zeros(n, dtype=w.dtype).index_add(0, d2, w2)
    dis = deg.pow(-0.5)
    dis = torch.where(torch.isinf(dis), torch.zeros_like(dis), dis)
    return s2, d2, dis[s2] * w2 * dis[d2]


def gcn_conv(x: Tensor, edge_index: Tensor, w: Optional[Tensor], weight: Tensor, bias: Optional[Tensor]):
    n = x.shape[0]
    if w is None:
        w = torch.ones(edge_index.shape[1], dtype=x.dtype)
    s2, d2, nrm = gcn_norm(edge_index, w, n)
    u = x @ weight.t()
    out = torch.zeros(n, u.shape[1], dtype=u.dtype).index_add(0, d2, nrm[:, None] * u[s2])
    return out if bias is None else out + bias


# a4. encoder stack: relu(conv) per layer, cat, per-graph view (kernel/sgcn_img_snp.py:218-228)
def _relu(z, pattern=None):
    """ReLU, or -- when `pattern` (a boolean tensor of z's shape) is given -- the same piecewise-linear map with the active set
    FIXED to `pattern`.  Test infrastructure for large inputs: an fp32 kernel and this fp64 oracle legitimately disagree on the
    sign of pre-activations within rounding error of zero (about one element per million), and each such flip moves every gradient
    that flows through it by a full term; with the kernel's own active set imposed, the comparison is about the arithmetic again.
    The tests also assert that the flipped elements are few and within rounding distance of zero."""
    if pattern is None:
        return torch.relu(z)
    return z * pattern.to(z.dtype)


def sgcn_encoder(P, x, edge_index, w, num_layers: int, rois: int, prefix: str = "", relu_pattern=None, pre_out=None):
    """relu_pattern: optional (N, L*H) boolean active set (see _relu); pre_out: optional list that receives the pre-activations."""
    hs, h = [], x
    for l in range(num_layers):
        name = "conv1" if l == 0 else "convs.%d" % (l - 1)
        z = gcn_conv(h, edge_index, w, P[prefix + name + ".lin.weight"], P[prefix + name + ".bias"])
        if pre_out is not None:
            pre_out.append(z.detach())
        H = z.shape[1]
        h = _relu(z, None if relu_pattern is None else relu_pattern.reshape(-1, relu_pattern.shape[-1])[:, l * H:(l + 1) * H])
        hs.append(h)
    cat = torch.cat(hs, 1)
    return cat.view(-1, rois, cat.shape[1])        # to_dense_batch is a view: every graph has `rois` nodes


# --------------------------------------------------------------------------------------
# a5. GATConv(in,out,edge_dim=1), heads=1  (PyG 2.0.2; call sites kernel/sgcn.py:163-166)
# --------------------------------------------------------------------------------------
def gat_conv(x, edge_index, edge_attr, lin_w, att_src, att_dst, lin_edge_w, att_edge, bias, slope=0.2):
    n = x.shape[0]
    h = x @ lin_w.t()
    a_s, a_d = h @ att_src.view(-1), h @ att_dst.view(-1)
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    src, dst, ea = src[keep], dst[keep], edge_attr[keep]
    cnt = torch.zeros(n, dtype=x.dtype).index_add(0, dst, torch.ones_like(ea))
    mean = torch.zeros(n, dtype=x.dtype).index_add(0, dst, ea) / cnt.clamp(min=1)   # fill_value='mean'
    loops = torch.arange(n)
    src, dst, ea = torch.cat([src, loops]), torch.cat([dst, loops]), torch.cat([ea, mean])
    ce = (lin_edge_w.view(-1) * att_edge.view(-1)).sum()          # Linear(1->C) then <., att_edge>
    logit = F.leaky_relu(a_s[src] + a_d[dst] + ea * ce, slope)
    mx = torch.full((n,), -float("inf"), dtype=x.dtype).scatter_reduce(0, dst, logit, "amax")
    ex = (logit - mx[dst].detach()).exp()
    den = torch.zeros(n, dtype=x.dtype).index_add(0, dst, ex)
    alpha = ex / (den[dst] + 1e-16)
    out = torch.zeros_like(h).index_add(0, dst, alpha[:, None] * h[src])
    return out + bias


# --------------------------------------------------------------------------------------
# a6. GO index preparation  (kernel/go_model.py:42-74, 161-168) -- integer work, bit exact
# --------------------------------------------------------------------------------------
def go_index_prep(A_dense: np.ndarray, Ag_dense: np.ndarray, pool: List[int], n_l: int = 2):
    """A_dense[parent, child] = 1 (G,G) ; Ag_dense (G,S).  Returns, per encoder layer j, the
    row-major nnz list of A[off_j:, off_j:] (go_model.py:52-60) and per decoder layer the nnz
    list of A^T[sum(pool[:n_l-j-1]):, sum(pool[:n_l-j]):] (go_model.py:69-73); `store` is the
    compact rank of each nnz's row among the non-empty rows (go_model.py:161-168)."""
    def nnz_rowmajor(M):
        r, c = np.nonzero(M)
        return np.stack([r, c]).astype(np.int64)

    def compact(rows):
        if rows.size == 0:
            return rows.copy()
        change = np.concatenate([[0], (rows[1:] != rows[:-1]).astype(np.int64)])
        return np.cumsum(change)

    enc, off = [], 0
    for j in range(n_l):
        off += 0 if j == 0 else pool[j - 1]
        ind = nnz_rowmajor(A_dense[off:, off:])
        enc.append(dict(index=ind, store=compact(ind[0]), nodes=A_dense.shape[0] - off))
    At, dec = A_dense.T, []
    for j in range(n_l):
        r0, c0 = sum(pool[: n_l - j - 1]), sum(pool[: n_l - j])
        ind = nnz_rowmajor(At[r0:, c0:])
        dec.append(dict(index=ind, store=compact(ind[0]), rows=At.shape[0] - r0, cols=At.shape[0] - c0))
    return dict(enc=enc, dec=dec, ag=nnz_rowmajor(Ag_dense), ag_t=nnz_rowmajor(Ag_dense.T),
                G=A_dense.shape[0], S=Ag_dense.shape[1], pool=list(pool), n_l=n_l)


# --------------------------------------------------------------------------------------
# a7-a10. GO network forward  (kernel/go_model.py:205-287)
# --------------------------------------------------------------------------------------
def _bn(P, name, x, training, stats_out=None):
    """BatchNorm1d on (B,C) or (B,C,L): batch statistics when training (biased var), else running."""
    dims = [0] if x.dim() == 2 else [0, 2]
    shape = [1, -1] if x.dim() == 2 else [1, -1, 1]
    if training:
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
        if stats_out is not None:
            stats_out[name] = (mean.detach(), var.detach(), x.numel() // x.shape[1])
    else:
        mean, var = P[name + ".running_mean"], P[name + ".running_var"]
    xh = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + 1e-5)
    return xh * P[name + ".weight"].view(shape) + P[name + ".bias"].view(shape)


def _ln_nodes(x, gamma, beta):
    """LayerNorm over the NODE axis of (B,M,d) per (b, channel) (go_model.py:246: permute, LN, permute)."""
    mu = x.mean(1, keepdim=True)
    var = x.var(1, unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5) * gamma.view(1, -1, 1) + beta.view(1, -1, 1)


def go_attention_layer(x, index, W_inc, W_s, u_att, v_s, per_subject_loop=False):
    """One encoder layer before LN (go_model.py:226-244).  index = (row, col) nnz of the layer's
    sub-adjacency; row aggregates from col."""
    row, col = torch.as_tensor(index[0]), torch.as_tensor(index[1])
    M, d = x.shape[1], W_inc.shape[0]
    xin, xs = x @ W_inc.t(), x @ W_s.t()
    a = torch.exp(torch.tanh(xin[:, row] @ u_att[0, :d] + xin[:, col] @ u_att[0, d:]))   # (B,nnz), row half first (:232)
    gate = torch.sigmoid(xs @ v_s[0])                                                  # (B,M)
    if per_subject_loop:
        # mirrors the reference's O(B) loop of sparse ops (go_model.py:236-244); used for cpu_baseline timing
        outs = []
        for k in range(x.shape[0]):
            s = torch.zeros(M, dtype=x.dtype).index_add(0, row, a[k])
            A_hat = torch.sparse_coo_tensor(torch.stack([row, col]), a[k] / s[row], (M, M))
            outs.append(torch.sparse.mm(A_hat, xin[k]) + xs[k] * gate[k][:, None])
        return torch.stack(outs)
    s = torch.zeros(x.shape[0], M, dtype=x.dtype).index_add(1, row, a)
    alpha = a / s[:, row]
    agg = torch.zeros(x.shape[0], M, d, dtype=x.dtype).index_add(1, row, alpha[:, :, None] * xin[:, col])
    return agg + xs * gate[:, :, None]


def go_decoder_layer(x, index, rows, self_off, W_out, W_s):
    """One decoder layer before LN (go_model.py:262-272): uniform 1/|row| weights, self term on rows >= self_off."""
    row, col = torch.as_tensor(index[0]), torch.as_tensor(index[1])
    xo, xs = x @ W_out.t(), x @ W_s.t()
    cnt = torch.zeros(rows, dtype=x.dtype).index_add(0, row, torch.ones(row.numel(), dtype=x.dtype))
    v = 1.0 / cnt[row]
    out = torch.zeros(x.shape[0], rows, xo.shape[2], dtype=x.dtype).index_add(1, row, v[None, :, None] * xo[:, col])
    pad = torch.zeros(x.shape[0], self_off, xo.shape[2], dtype=x.dtype)
    return out + torch.cat([pad, xs], 1)


GO_MASK_NAMES = ["go_enc0", "go_enc1", "go_B", "go_dec0", "go_dec1", "go_BD", "go_latent"]


def go_forward(P, prep, data, training: bool, masks: Optional[Dict[str, Tensor]] = None,
               prefix: str = "go_network.", per_subject_loop=False, stats_out=None):
    """Gene_ontology_network.forward.  `masks[name]` are multiplicative dropout scale tensors
    (0 or 1/(1-p)), in the order the reference draws them; None = no dropout (eval)."""
    p = lambda n: P[prefix + n]
    mk = (lambda n, t: t * masks[n]) if (training and masks is not None) else (lambda n, t: t)
    G, S, pool, n_l = prep["G"], prep["S"], prep["pool"], prep["n_l"]
    gi, si = torch.as_tensor(prep["ag"][0]), torch.as_tensor(prep["ag"][1])
    B = data.shape[0]
    # a7 encode (go_model.py:208-215): x[b,g,c] = sum_s t_c[(g,s)] data[b,s]
    chans = []
    for c in range(2):
        chans.append(torch.zeros(B, G, dtype=data.dtype).index_add(1, gi, data[:, si] * p("t.%d" % c)[None, :]))
    x = torch.stack(chans, 2)
    # a8 encoder layers (go_model.py:219-251)
    for j in range(n_l):
        e = prep["enc"][j]
        o = go_attention_layer(x, e["index"], p("w_inc.%d.weight" % j), p("w_s_loop.%d.weight" % j),
                               p("w_att_in.%d.weight" % j), p("w_att_s.%d.weight" % j), per_subject_loop)
        o = torch.relu(_ln_nodes(o, p("G_B.%d.weight" % j), p("G_B.%d.bias" % j)))
        o = mk("go_enc%d" % j, o)
        x = o[:, pool[j]:, :]
    # a10 read-outs (go_model.py:254-255)
    att = x @ p("conc_for_attention.0.weight").t()
    atten_out = torch.relu(_bn(P, prefix + "conc_for_attention.1", att, training, stats_out))
    inp = (x @ p("conc.weight").t()).squeeze(-1)
    inp_out = mk("go_B", torch.relu(_bn(P, prefix + "B.0", inp, training, stats_out)))
    # a9 decoder (go_model.py:258-275)
    for j in range(n_l):
        d = prep["dec"][j]
        o = go_decoder_layer(x, d["index"], d["rows"], pool[n_l - j - 1],
                             p("w_out.%d.weight" % j), p("w_s_loop_out.%d.weight" % j))
        o = torch.relu(_ln_nodes(o, p("G_B_D.%d.weight" % j), p("G_B_D.%d.bias" % j)))
        x = mk("go_dec%d" % j, o)
    out_D = (x @ p("conc_D.weight").t()).squeeze(-1)
    out_D = mk("go_BD", torch.relu(_bn(P, prefix + "B_D.0", out_D, training, stats_out)))
    # a7 decode (go_model.py:281-282): x_D[b,s] = sum_g t_D[(s,g)] out_D[b,g]
    ri, ci = torch.as_tensor(prep["ag_t"][0]), torch.as_tensor(prep["ag_t"][1])
    x_D = torch.zeros(B, S, dtype=data.dtype).index_add(1, ri, out_D[:, ci] * p("t_D.0")[None, :])
    # latent MLP (go_model.py:138-146, 285)
    h = inp_out @ p("latent.0.weight").t()
    h = mk("go_latent", torch.relu(_bn(P, prefix + "latent.1", h, training, stats_out)))
    h = h @ p("latent.4.weight").t()
    latent = torch.relu(_bn(P, prefix + "latent.5", h, training, stats_out))
    return latent, x_D, atten_out


# --------------------------------------------------------------------------------------
# a11. cross attention + fusion heads ; full model forward (kernel/sgcn_img_snp.py:207-307)
# --------------------------------------------------------------------------------------
def cross_attention(P, q, kv, heads=2, prefix="multihead_attn."):
    """nn.MultiheadAttention(E, 2, batch_first=True)(q, kv, kv)[0], dropout 0 (:46,:240)."""
    E = q.shape[-1]
    W, b = P[prefix + "in_proj_weight"], P[prefix + "in_proj_bias"]
    Q = q @ W[:E].t() + b[:E]
    K = kv @ W[E:2 * E].t() + b[E:2 * E]
    V = kv @ W[2 * E:].t() + b[2 * E:]
    B, Lq, Lk, hd = q.shape[0], q.shape[1], kv.shape[1], E // heads
    split = lambda t, L: t.view(B, L, heads, hd).transpose(1, 2)
    s = (split(Q, Lq) @ split(K, Lk).transpose(-1, -2)) / math.sqrt(hd)
    o = (torch.softmax(s, -1) @ split(V, Lk)).transpose(1, 2).reshape(B, Lq, E)
    return o @ P[prefix + "out_proj.weight"].t() + P[prefix + "out_proj.bias"]


MODEL_MASK_NAMES = GO_MASK_NAMES + ["lin1", "lin1_regr"]


def model_forward(P, prep, batch: Dict[str, Tensor], num_layers: int, rois: int, explain: bool,
                  training: bool, masks=None, per_subject_loop=False, stats_out=None, patterns=None, probe=None):
    """SGCN_GCN_IMGSNP.forward with isCrossAtten=True, isuseProb4Regr=True, image+SNP fusion
    (the default configuration, main.py:49,65).  Returns the reference's 6-tuple.
    patterns: optional dict {'enc': (B,R,LH) bool, 'attn': (B,R,LH) bool} of imposed ReLU active sets (see _relu);
    probe: optional dict that receives the pre-activations of those two ReLU sites."""
    x, ei, w, snps = batch["x"], batch["edge_index"], batch["edge_attr"], batch["snps_feat"]
    mk = (lambda n, t: t * masks[n]) if (training and masks is not None) else (lambda n, t: t)
    if explain:
        m = cal_probability(P, x, ei, w, rois, snps)
        xe, we, se = m["x"], m["w"], m["snps"]
    else:
        xe, we, se = x, w, snps
    pre = [] if probe is not None else None
    batch_x = sgcn_encoder(P, xe, ei, we, num_layers, rois, relu_pattern=None if patterns is None else patterns["enc"], pre_out=pre)
    B = batch_x.shape[0]
    img_out = batch_x.reshape(B, -1)
    latent, x_hat, atten_out = go_forward(P, prep, se, training, masks, per_subject_loop=per_subject_loop,
                                          stats_out=stats_out)
    attn_pre = cross_attention(P, batch_x, atten_out)
    if probe is not None:
        probe["enc"] = torch.cat(pre, 1).view(B, rois, -1)
        probe["attn"] = attn_pre.detach()
    out_cross = _relu(attn_pre, None if patterns is None else patterns["attn"]).reshape(B, -1)
    out_z = (img_out + out_cross) / 2
    out_lin = torch.cat([out_z, latent], -1)
    linear_outf = torch.relu(out_lin @ P["lin1.weight"].t() + P["lin1.bias"])
    logits = mk("lin1", linear_outf) @ P["lin2.weight"].t() + P["lin2.bias"]
    img_feat = (x.view(B, rois, -1) * P["prob"]).reshape(B, -1)         # data.x (unmasked) * prob (:293-297)
    reg = torch.relu(torch.cat([out_lin, img_feat], -1) @ P["lin1_regr.weight"].t() + P["lin1_regr.bias"])
    reg = mk("lin1_regr", reg) @ P["lin2_regr.weight"].t() + P["lin2_regr.bias"]
    return torch.log_softmax(logits, -1), x_hat, out_z, out_lin, linear_outf, reg


# --------------------------------------------------------------------------------------
# a12/a13. losses  (kernel/sgcn_img_snp.py:153-205) and the train() step body (train_eval...:516-545)
# --------------------------------------------------------------------------------------
LAMDA = dict(x_l1=0.1, e_l1=0.1, x_ent=0.1, e_ent=0.1, mi=1.0, ce=1.0)      # sgcn_hyperparameters.py:18-23


def _l1_ent(p, eps=1e-6):
    n = p.numel()
    return p.abs().sum() / n, -(p * torch.log(p + eps) + (1 - p) * torch.log(1 - p + eps)).sum() / n


def loss_probability(P, x, edge_index, edge_weight, rois, with_snps=True):
    pe = cal_probability(P, x, edge_index, edge_weight, rois)["p_e"]
    f_l1, f_en = _l1_ent(torch.sigmoid(P["prob"]))
    e_l1, e_en = _l1_ent(pe)
    l1 = LAMDA["x_l1"] * f_l1 + LAMDA["e_l1"] * e_l1
    en = LAMDA["x_ent"] * f_en + LAMDA["e_ent"] * e_en
    if with_snps:
        s_l1, s_en = _l1_ent(torch.sigmoid(P["snps_prob"]))
        l1 = l1 + LAMDA["x_l1"] * s_l1
        en = en + LAMDA["x_ent"] * s_en
    return l1 + en


def consist_loss(s, tsne=None, gamma=0.005):
    n = s.shape[0]
    if n == 0:
        return s.new_zeros(())
    W = torch.exp(-gamma * torch.cdist(tsne, tsne, p=2) ** 2) if tsne is not None else torch.ones(n, n, dtype=s.dtype)
    L = torch.diag(W.sum(1)) - W
    return torch.trace(s.t() @ L @ s) / (n * n)


def orthogonal_constraint(w):
    wn = w / w.norm(dim=1)[:, None]
    return torch.norm(wn.t() @ wn - torch.eye(wn.shape[1], dtype=w.dtype)) ** 2 / (wn.shape[0] ** 2)


def train_step_loss(P, prep, batch, num_layers, rois, lambda_loss, rbf_gamma, training=True,
                    masks_plain=None, masks_explain=None, per_subject_loop=False, with_orth=True, patterns=None, probes=None):
    """The scalar that train() back-propagates (train_eval_sgcn_img_snps.py:521-544), isSoftSimilarity=True.
    patterns / probes: optional (plain, explain) pairs passed to model_forward."""
    y, cs, snps = batch["y"], batch["clini_score"], batch["snps_feat"]
    pp, pe_ = (None, None) if patterns is None else patterns
    bp, be_ = (None, None) if probes is None else probes
    o = model_forward(P, prep, batch, num_layers, rois, False, training, masks_plain, per_subject_loop, patterns=pp, probe=bp)
    q = model_forward(P, prep, batch, num_layers, rois, True, training, masks_explain, per_subject_loop, patterns=pe_, probe=be_)
    lam = lambda_loss
    loss_ce = lam[0] * F.nll_loss(o[0], y)
    loss_mi = lam[0] * F.nll_loss(q[0], y)
    loss_reg = lam[1] * (F.mse_loss(o[5].reshape(-1), cs) + F.mse_loss(q[5].reshape(-1), cs)) / 2
    loss_prob = lam[2] * loss_probability(P, batch["x"], batch["edge_index"], batch["edge_attr"], rois)
    recon = lam[3] * (((o[1] - snps) ** 2).sum() + ((q[1] - snps) ** 2).sum()) / 2
    clus = lam[4] * (consist_loss(o[2], batch["tsne_fdim"], rbf_gamma) + consist_loss(q[2], batch["tsne_fdim"], rbf_gamma)) / 2
    orth = lam[5] * orthogonal_constraint(o[2]) if with_orth else 0.0
    if lam[0] == 0:
        loss_ce = loss_mi = 0.0
    total = LAMDA["ce"] * loss_ce + LAMDA["mi"] * loss_mi + loss_reg + loss_prob + recon + clus + orth
    return total, o, q


# --------------------------------------------------------------------------------------
# config 1: image-only SGCN_GCN (kernel/sgcn.py:272-388) and its 3-term step (train_eval_sgcn.py:296-313)
# --------------------------------------------------------------------------------------
def sgcn_gcn_forward(P, batch, num_layers, rois, explain, training, mask=None):
    x, ei, w = batch["x"], batch["edge_index"], batch["edge_attr"]
    if explain:
        m = cal_probability(P, x, ei, w, rois)
        x, w = m["x"], m["w"]
    z = sgcn_encoder(P, x, ei, w, num_layers, rois).reshape(-1, rois * num_layers * P["conv1.bias"].numel())
    h = torch.relu(z @ P["lin1.weight"].t() + P["lin1.bias"])
    if training and mask is not None:
        h = h * mask
    return torch.log_softmax(h @ P["lin2.weight"].t() + P["lin2.bias"], -1)


def loss_probability_sgcn(P, x, edge_index, edge_weight, rois):
    """kernel/sgcn.py:329-351 variant: per-row L1 / N for the node mask, no SNP term."""
    pe = cal_probability(P, x, edge_index, edge_weight, rois)["p_e"]
    xp = torch.sigmoid(P["prob"])
    f_l1 = xp.abs().sum(-1).sum() / xp.shape[0]
    _, f_en = _l1_ent(xp)
    e_l1, e_en = _l1_ent(pe)
    return LAMDA["x_l1"] * f_l1 + LAMDA["e_l1"] * e_l1 + LAMDA["x_ent"] * f_en + LAMDA["e_ent"] * e_en
