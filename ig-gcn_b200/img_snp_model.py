"""`SGCN_GCN_IMGSNP` on the fused kernels (reference: kernel/sgcn_img_snp.py).

Constructor arguments, forward signature `model(data, temperature, device, isExplain=False)`, the 6-tuple it
returns, the helper methods the training loop calls (`loss_probability`, `consist_loss`,
`OrthogonalConstraint`, `cal_probability`) and every parameter name/shape follow the reference, so
kernel/train_eval_sgcn_img_snps.py::train() runs against it unchanged and checkpoints interchange.

What changed inside:
  * mask -> norm -> L x GCNConv -> relu -> cat -> to_dense_batch is ONE kernel (ops.sgcn_encoder); the two
    `x.min().item()` host syncs (sgcn_img_snp.py:225,293) disappear because every graph has `rois` nodes,
    so to_dense_batch is a view;
  * loss_probability reuses the p_e computed by the explain pass instead of a third cal_probability;
  * consist_loss / OrthogonalConstraint use the B x B Gram form (no D x D intermediates);
  * `n_snps` comes from A_g (the reference hard-codes 54, sgcn_img_snp.py:96).
"""
from __future__ import annotations

import math
import os
import weakref

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import init
from torch.nn.parameter import Parameter

from . import ops
from .data import Batch
from .go_net import Gene_ontology_network
from .pyg import GCNConv            # PyG-signature operator; here the holder of `conv*.lin.weight` / `conv*.bias` (the stack runs fused)


# IGCN_ONE_STREAM=1 keeps the whole step on one stream (A/B hook)
_TWO_STREAMS = os.environ.get("IGCN_ONE_STREAM", "") == ""
_PREFETCH_CONSIST = os.environ.get("IGCN_NO_PREFETCH_CONSIST", "") == ""
if _TWO_STREAMS and hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
    # parameters shared by nodes on different streams make autograd warn about the AccumulateGrad stream; the extra event wait it
    # mentions is intended here
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)


def get_csr(data):
    csr = getattr(data, "csr", None) if isinstance(data, Batch) else getattr(data, "_igcn_csr", None)
    return csr


class MaskedEncoderMixin:
    """cal_probability / loss bookkeeping shared by the SGCN model family."""

    def _csr_for(self, data, rois):
        # Called first by every forward: a new forward may be a new batch, and `data.to(device)` of a fixed-size loader tends to
        # land on the addresses the previous batch just freed, so nothing keyed on a device address may outlive the forward that
        # made it.  The intended reuse survives: the explain forward of a step stores p_e for that step's loss_probability, and
        # the first consist_loss of a step stores the similarity matrix for the second one.
        self._pe_cache = None
        self._w_cache = None
        csr = get_csr(data)
        if csr is None:
            b = Batch.from_device_tensors(data.x.detach(), data.edge_index, data.edge_attr, rois)
            csr = b.csr
            try:
                data._igcn_csr = csr
            except Exception:
                pass
        self._last_csr = (data.edge_index.data_ptr(), data.edge_index.shape[1], csr, weakref.ref(data.edge_index))
        return csr

    def _csr_lookup(self, x, edge_index, edge_weight):
        last = getattr(self, "_last_csr", None)
        # valid only while the edge_index of the last forward is alive (its address cannot have been recycled then)
        if last is not None and last[3]() is not None and last[0] == edge_index.data_ptr() and last[1] == edge_index.shape[1]:
            return last[2]
        return Batch.from_device_tensors(x.detach(), edge_index, edge_weight, self.rois).csr

    def _conv_params(self):
        convs = [self.conv1] + list(self.convs)
        return [c.lin.weight for c in convs], [c.bias for c in convs]

    def _edge_prob(self, x, edge_index, edge_weight):
        """p_e in CSR-slot order; reuses the explain pass's value when it was computed on the same inputs."""
        c = getattr(self, "_pe_cache", None)
        key = (x.data_ptr(), edge_index.data_ptr(), self.prob._version, self.prob_bias._version)
        if c is not None and c[0] == key and torch.is_grad_enabled() == c[2]:
            return c[1]
        csr = self._csr_lookup(x, edge_index, edge_weight)
        return ops.edge_mask(x, csr, self.prob, self.prob_bias)

    def cal_probability(self, x, edge_index, edge_weight, snps_feat=None):
        """Reference surface (sgcn_img_snp.py:133-151).  edge outputs are returned in ORIGINAL edge order."""
        csr = self._csr_lookup(x, edge_index, edge_weight)
        N, D = x.shape
        x_feat_prob = (x.view(N // self.rois, self.rois, D) * self.prob).reshape(N, D)
        pe_csr = ops.edge_mask(x, csr, self.prob, self.prob_bias)
        edge_prob = torch.empty_like(pe_csr).index_copy(0, csr.csr_perm.long(), pe_csr)
        edge_weight_prob = edge_weight * edge_prob
        if snps_feat is not None:
            sp = torch.sigmoid(self.snps_prob)
            return x_feat_prob, edge_weight_prob, self.prob, edge_prob, snps_feat * sp, sp
        return x_feat_prob, edge_weight_prob, self.prob, edge_prob


def _l1_entropy(p, eps=1e-6):
    n = p.numel()
    return p.abs().sum() / n, -(p * torch.log(p + eps) + (1 - p) * torch.log(1 - p + eps)).sum() / n


class SGCN_GCN_IMGSNP(nn.Module, MaskedEncoderMixin):

    def __init__(self, num_layers, hidden, A_g, A, pool_dim, l_dim, device, *args, hidden_linear=64, rois=90, H_0=3,
                 num_classes=2, isCrossAtten=False, isSoftSimilarity=False, rbf_gamma=0.005, graph_pool=False,
                 isuseProb4Regr=False, num_regr=4, model4eachregr=False, isImageOnly=True, isSNPsOnly=False,
                 isMultiFusion=False, **kwargs):
        super().__init__()
        self.device = device
        self.isCrossAtten, self.isSoftSimilarity, self.rbf_gamma = isCrossAtten, isSoftSimilarity, rbf_gamma
        self.model4eachregr, self.isuseProb4Regr = model4eachregr, isuseProb4Regr
        self.isImageOnly, self.isSNPsOnly, self.num_regr = isImageOnly, isSNPsOnly, num_regr
        self.input = None
        self.final_conv_acts = None
        self.final_conv_grads = None
        self.rois, self.prob_dim, self.isMultiFusion = rois, H_0, isMultiFusion
        self.num_layers, self.hidden = num_layers, hidden
        n_snps = int(A_g.shape[1])
        self.n_snps = n_snps
        self.conv1 = GCNConv(H_0, hidden)
        self.convs = nn.ModuleList()
        n_l = 2
        dim_snps_atten = hidden
        if isCrossAtten:
            for _ in range(num_layers - 2):
                self.convs.append(GCNConv(hidden, hidden))
                dim_snps_atten += hidden
            self.convs.append(GCNConv(hidden, hidden))
            dim_snps_atten += hidden
            self.pool = pool_dim[0]
            self.multihead_attn = nn.MultiheadAttention(dim_snps_atten, 2, batch_first=True)
        else:
            for _ in range(num_layers - 1):
                self.convs.append(GCNConv(hidden, hidden))
        self.graph_pool = graph_pool
        enc_dim = rois * (1 + len(self.convs)) * hidden
        if graph_pool:
            self.lin1 = nn.Linear(3 * num_layers * hidden + l_dim, hidden_linear)
            self.lin1_regr = nn.Linear(3 * num_layers * hidden + l_dim, hidden_linear)
        else:
            if isImageOnly:
                self.lin1 = nn.Linear(rois * num_layers * hidden, hidden_linear)
            elif isSNPsOnly:
                self.lin1 = nn.Linear(l_dim + n_snps, hidden_linear)
            else:
                self.lin1 = nn.Linear(rois * num_layers * hidden + l_dim, hidden_linear)
            extra = rois * H_0 if isuseProb4Regr else 0
            if isImageOnly:
                self.lin1_regr = nn.Linear(rois * num_layers * hidden + extra, hidden_linear)
            elif isSNPsOnly:
                self.lin1_regr = nn.Linear(l_dim + n_snps, hidden_linear)
            else:
                self.lin1_regr = nn.Linear(rois * num_layers * hidden + l_dim + extra, hidden_linear)
        self.lin2 = nn.Linear(hidden_linear, num_classes)
        self.lin2_regr = nn.Linear(hidden_linear, num_regr)
        self.batch_norm_1d = nn.BatchNorm1d(num_features=rois * num_layers * hidden + l_dim)     # unused, kept for state_dict parity
        self.prob = Parameter(torch.empty((rois, H_0)))
        self.prob_bias = Parameter(torch.empty((H_0 * 2, 1)))
        init.kaiming_uniform_(self.prob_bias, a=math.sqrt(5))
        self.edge_prob = Parameter(torch.empty((rois, rois)))                                      # unused by the reference too
        init.kaiming_uniform_(self.prob, a=math.sqrt(5))
        init.kaiming_uniform_(self.edge_prob, a=math.sqrt(5))
        self.snps_prob = Parameter(torch.empty((1, n_snps)))
        init.kaiming_uniform_(self.snps_prob, a=math.sqrt(5))
        self.go_network = Gene_ontology_network(A_g, A, 2, n_l, [5, 5], pool_dim, l_dim, device, dim_snps_atten=dim_snps_atten)
        self.batch_norm = nn.BatchNorm1d(num_layers * hidden)                                      # unused, state_dict parity
        self.dropout_masks = None     # test hook: dict name -> scale tensor (oracle.MODEL_MASK_NAMES)
        self._enc_dim = enc_dim

    def reset_parameters(self):
        self.conv1.reset_parameters()
        for conv in self.convs:
            conv.reset_parameters()
        for m in (self.lin1, self.lin2, self.lin1_regr, self.lin2_regr):
            m.reset_parameters()
        with torch.no_grad():
            init.kaiming_uniform_(self.prob_bias, a=math.sqrt(5))
            init.kaiming_uniform_(self.prob, a=math.sqrt(5))
            init.kaiming_uniform_(self.edge_prob, a=math.sqrt(5))
            init.kaiming_uniform_(self.snps_prob, a=math.sqrt(5))

    def activations_hook(self, grad):
        self.final_conv_grads = grad

    # --------------------------------------------------------------------------------------------------
    def loss_probability(self, x, edge_index, edge_weight, hp, eps=1e-6):
        """sgcn_img_snp.py:153-181 (sums over edges are order independent, so CSR-slot order is used as is)."""
        edge_prob = self._edge_prob(x, edge_index, edge_weight)
        return ops.mask_loss(self.prob, edge_prob, self.snps_prob, hp, eps)        # one fused reduction (glue.cu); CUDA only

    def _similarity(self, n, tsne_result, like):
        """(W, d): the similarity matrix of the batch (RBF of tsne_fdim; None = all ones) and its row sums; depends on the data
        only, so the two passes of a step share it."""
        soft = self.isSoftSimilarity and tsne_result is not None
        if not soft:
            return None, None
        key = (tsne_result.data_ptr(), tsne_result._version, n)
        c = getattr(self, "_w_cache", None)
        if c is not None and c[0] == key:
            return c[1]
        wd = ops.rbf_similarity(tsne_result, self.rbf_gamma)
        self._w_cache = (key, wd)
        return wd

    def consist_loss(self, s, tsne_result=None):
        """tr(s^T (D-W) s)/B^2 (sgcn_img_snp.py:183-196).  With L = D - W this is <s, L s>/B^2: one (B x B)(B x D) product gives the
        value (a dot product) and the gradient 2 L s / B^2 -- the reference's D x D intermediates (8 448^2 at 264 ROIs) and a second
        product in the backward never exist; ops.laplacian_quadratic centres the columns first so the product does not cancel."""
        n = s.shape[0]
        if n == 0:
            return 0
        W, d = self._similarity(n, tsne_result, s)
        return ops.laplacian_quadratic(s, W, d, 1.0 / (n * n))

    def consist_loss_pair(self, s2, tsne_result=None):
        """consist_loss(s2[:B], t) + consist_loss(s2[B:], t) for the stacked plain / explain features of forward_pair: one
        product and one dot product for both passes, and no slice in the autograd graph."""
        c = getattr(self, "_quad_cache", None)
        if c is not None:
            self._quad_cache = None
            if c[0] is s2 and c[1] is tsne_result:       # already queued on the third stream by forward_pair(consist=True)
                cur = torch.cuda.current_stream(s2.device)
                cur.wait_stream(c[3])
                c[2].record_stream(cur)
                return c[2]
        n = s2.shape[0] // 2
        W, d = self._similarity(n, tsne_result, s2)
        return ops.laplacian_quadratic(s2, W, d, 1.0 / (n * n), halves=2)

    def OrthogonalConstraint(self, w):
        """||w^T w - I_D||_F^2 / B^2 with row-normalised w (sgcn_img_snp.py:198-205) = (||w w^T||_F^2 - 2B + D)/B^2: the B x B Gram
        matrix instead of the reference's D x D product (8 448^2 at 264 ROIs), on the tensor cores (ops.gram)."""
        wn = w / w.norm(dim=1)[:, None]
        g = ops.gram(wn)
        n, d = wn.shape
        return ((g * g).sum() - 2.0 * g.diagonal().sum() + d) / (n * n)

    # --------------------------------------------------------------------------------------------------
    def _mask(self, name, t, p):
        if not self.training:
            return t
        if self.dropout_masks is not None:
            return t * self.dropout_masks[name].to(t.device).float()
        return t * self.go_network.mask_bank.get(name, t.shape, p)

    def _mask_of(self, name, t, p):
        """the multiplicative dropout scale `_mask` would apply (None in eval mode)"""
        if not self.training:
            return None
        if self.dropout_masks is not None:
            return self.dropout_masks[name].to(t.device).float().expand_as(t)
        return self.go_network.mask_bank.get(name, t.shape, p)

    def forward(self, data, temperature=None, device=None, isExplain=False):
        x, edge_index, edge_weight = data.x, data.edge_index, data.edge_attr
        snps_feat = data.snps_feat
        if not x.requires_grad and x.is_leaf:
            x.requires_grad = True                     # the reference wants dLoss/dx (sgcn_img_snp.py:210)
        self.input = x
        csr = self._csr_for(data, self.rois)
        Ws, bs = self._conv_params()
        bank = self.go_network.mask_bank
        use_bank = self.training and self.dropout_masks is None
        if use_bank:
            bank.begin_pass(csr.B, x.device)           # all nine dropout masks of this pass: one launch
        if isExplain:
            batch_x, p_e = ops.sgcn_encoder(x, csr, Ws, bs, self.prob, self.prob_bias, want_pe=True)
            self._pe_cache = ((x.data_ptr(), edge_index.data_ptr(), self.prob._version, self.prob_bias._version), p_e,
                              torch.is_grad_enabled())
            snps_feat_prob = snps_feat * torch.sigmoid(self.snps_prob)
        else:
            batch_x, _ = ops.sgcn_encoder(x, csr, Ws, bs)
            snps_feat_prob = snps_feat
        B = batch_x.shape[0]
        img_out = batch_x.view(B, -1)
        if self.graph_pool:
            img_out = torch.cat([batch_x.mean(1), batch_x.max(1)[0], batch_x.sum(1)], 1)
        go = self.go_network
        go.dropout_masks = self.dropout_masks
        latent, x_hat, _, atten_out = go(snps_feat_prob, temperature, device)
        fused_avg = self.isCrossAtten and not self.graph_pool and not self.isImageOnly and not self.isSNPsOnly
        if fused_avg:
            # out_z = (img_out + relu(MHA(q = ROI tokens, k = v = GO tokens))) / 2 as one kernel each way (cross_attn_mma.cuh)
            out_cross = None
        elif self.isCrossAtten:
            # relu(MHA(q = ROI tokens, k = v = GO tokens)) as one kernel (cross_attn.cu)
            out_cross = ops.cross_attention(batch_x, atten_out, self.multihead_attn, relu=True)
        else:
            out_cross = torch.cat((img_out, latent), -1)
        if out_cross is not None:
            if self.graph_pool:
                out_cross = torch.cat([out_cross.mean(1), out_cross.max(1)[0], out_cross.sum(1)], 1)
            else:
                out_cross = out_cross.reshape(B, -1)

        if self.isImageOnly:
            out_z = img_out
            parts = [out_z]
        elif self.isSNPsOnly:
            out_z = latent
            parts = [snps_feat_prob, latent]
        else:
            out_z = ops.cross_attention_average(batch_x, atten_out, self.multihead_attn).reshape(B, -1) if fused_avg \
                else (img_out + out_cross) / 2
            parts = [out_z, latent]
        # out_lin is part of the returned tuple (eval_scores collects it, train_eval...:626); the heads read its parts in place
        out_lin = parts[0] if len(parts) == 1 else torch.cat(parts, -1).detach()
        linear_outf = ops.cat_linear(parts, self.lin1.weight, self.lin1.bias, relu=True)
        rparts = parts
        if self.isuseProb4Regr and not self.isSNPsOnly:
            img_feat = (x.view(B, self.rois, -1) * self.prob).reshape(B, -1)       # data.x (unmasked) * prob (:293-297)
            rparts = parts + [img_feat]
        # relu(lin1_regr(cat(parts))) without building the concatenation; then dropout, lin2 + log_softmax and lin2_regr of both
        # heads in ONE launch (glue.cu), in training and in eval mode (masks None)
        r = ops.cat_linear(rparts, self.lin1_regr.weight, self.lin1_regr.bias, relu=True)
        logp, our_reg = ops.output_heads(linear_outf, self._mask_of("lin1", linear_outf, 0.5), r, self._mask_of("lin1_regr", r, 0.3),
                                         self.lin2, self.lin2_regr)
        if use_bank:
            bank.end_pass()
        return logp, x_hat, out_z, out_lin, linear_outf, our_reg

    def forward_pair(self, data, temperature=None, device=None, stacked=False, consist=False, mask_loss_hp=None):
        """Plain pass and explain pass of ONE batch in a single sweep: everything downstream of the two encoder
        launches (GO network, cross attention, fusion heads) runs once on the 2B stacked samples, with BatchNorm
        applied per pass, so the results equal `forward(data)` followed by `forward(data, isExplain=True)` while every
        parameter is used once (no gradient-accumulation kernels) and half as many kernels are launched.
        Returns (plain 6-tuple, explain 6-tuple).  Default configuration only (cross attention, image + SNP fusion)."""
        if not self.supports_pair():
            if stacked:
                raise RuntimeError("forward_pair(stacked=True) needs the default configuration (see supports_pair())")
            return self.forward(data, temperature, device), self.forward(data, temperature, device, isExplain=True)
        x, edge_index = data.x, data.edge_index
        snps = data.snps_feat
        if not x.requires_grad and x.is_leaf:
            x.requires_grad = True
        self.input = x
        csr = self._csr_for(data, self.rois)
        Ws, bs = self._conv_params()
        B = csr.B
        bank = self.go_network.mask_bank
        use_bank = self.training and self.dropout_masks is None
        if use_bank:
            bank.begin_pass(("pair", B), x.device)
        # The GO network depends only on the SNPs, the SGCN encoders only on the brain graphs: they run on two streams (fork /
        # join; inside a captured CUDA graph this becomes two parallel branches).  autograd replays every node on the stream of its
        # forward, so the GO backward also overlaps the attention / encoder backward.
        main = torch.cuda.current_stream(x.device)
        side = self._go_stream(x.device) if _TWO_STREAMS else None
        snps2 = ops.snp_mask_pair(snps, self.snps_prob)                             # [snps ; snps * sigmoid(snps_prob)]
        go = self.go_network
        go.dropout_masks = self.dropout_masks
        early = side is not None and os.environ.get("IGCN_EARLY_ATTN", "1") == "1"
        if side is not None:
            side.wait_stream(main)
            go.atten_ready = torch.cuda.Event() if early else None
            branch = self._go_stream(x.device, 2) if (early and os.environ.get("IGCN_GO_BRANCH", "1") == "1") else None
            go.branch_stream = branch                    # decoder branch beside the latent read-outs (go_net.forward)
            lat_st = self._go_stream(x.device, 3) if (branch is not None and os.environ.get("IGCN_GO_LATENT_STREAM", "1") == "1") else None
            go.latent_stream = lat_st
            with torch.cuda.stream(side):
                latent, x_hat, _, atten_out = go(snps2, temperature, device, groups=2)
            snps2.record_stream(side)
            atten_ev, go.atten_ready = go.atten_ready, None
            go.branch_stream = go.latent_stream = None
        side2 = self._go_stream(x.device, 1) if _TWO_STREAMS and os.environ.get("IGCN_ENC_STREAM", "1") == "1" else None
        # both passes write straight into the halves of ONE (2B, R, L*H) buffer: no torch.cat afterwards (277 MB each way at config 4)
        LH = sum(w.shape[0] for w in Ws)
        stacked_buf = torch.empty((2 * B, self.rois, LH), dtype=torch.float32, device=x.device)
        img_feat = None
        if side2 is not None:                                                       # the plain pass on a third stream
            side2.wait_stream(main)
            stacked_buf.record_stream(side2)
            with torch.cuda.stream(side2):
                h_plain, _ = ops.sgcn_encoder(x, csr, Ws, bs, out=stacked_buf[:B])
                if self.isuseProb4Regr:
                    # the regression head's masked image features depend on x and the node mask only: made here so that their
                    # backward (an elementwise product and a reduction over the batch) also runs beside the main stream
                    img_feat = (x.view(B, self.rois, -1) * self.prob).reshape(B, -1)
                plain_done = torch.cuda.Event()
                plain_done.record(side2)          # main joins HERE: the similarity matrix below is not needed before the consistency loss
                if consist and stacked and self.isSoftSimilarity and _PREFETCH_CONSIST:
                    # the similarity matrix of the consistency loss depends on the batch only: built here, off the path that
                    # later waits for out_z (consist_loss_pair finds it in the cache)
                    self._similarity(B, data.tsne_fdim, x)
        else:
            h_plain, _ = ops.sgcn_encoder(x, csr, Ws, bs, out=stacked_buf[:B])
        h_expl, p_e = ops.sgcn_encoder(x, csr, Ws, bs, self.prob, self.prob_bias, want_pe=True, out=stacked_buf[B:])
        self._pe_cache = ((x.data_ptr(), edge_index.data_ptr(), self.prob._version, self.prob_bias._version), p_e,
                          torch.is_grad_enabled())
        if side2 is not None:
            main.wait_event(plain_done)
            h_plain.record_stream(main)
            if img_feat is not None:
                img_feat.record_stream(main)
        if side2 is not None and mask_loss_hp is not None and stacked:
            # the mask loss needs only the explain encoder's edge probabilities: it (and its backward) runs on the third stream while
            # the attention occupies the main one; train.step_loss finds it in the cache
            ev = torch.cuda.Event()
            ev.record(main)
            side2.wait_event(ev)
            with torch.cuda.stream(side2):
                lp = self.loss_probability(x, edge_index, data.edge_attr, mask_loss_hp)
            p_e.record_stream(side2)
            self._lp_cache = (mask_loss_hp, lp, side2)
        batch_x = ops.join_halves(h_plain, h_expl, stacked_buf)                     # (2B, R, LH), no copy
        if side is not None:
            if early:
                main.wait_event(atten_ev)        # the attention needs only atten_out; decoder and latent MLP keep running on `side`
                atten_out.record_stream(main)
            else:
                main.wait_stream(side)
                for t in (latent, x_hat, atten_out):
                    t.record_stream(main)
        else:
            latent, x_hat, _, atten_out = go(snps2, temperature, device, groups=2)
        # out_z = (img_out + relu(attention)) / 2 in ONE kernel each way (the average is the attention kernels' epilogue)
        out_z = ops.cross_attention_average(batch_x, atten_out, self.multihead_attn).reshape(2 * B, -1)
        if consist and stacked and side2 is not None and self.isSoftSimilarity and _PREFETCH_CONSIST:
            # the consistency loss needs only out_z: its Laplacian product starts now on the third stream, beside the fusion heads
            tsne = data.tsne_fdim
            side2.wait_stream(main)
            with torch.cuda.stream(side2):
                quad = self.consist_loss_pair(out_z, tsne)
            out_z.record_stream(side2)
            self._quad_cache = (out_z, tsne, quad, side2)
        if side is not None and early:
            main.wait_stream(side)               # latent (fusion heads) is needed from here on
            if lat_st is not None:
                main.wait_stream(lat_st)
            latent.record_stream(main)
            if go.decoder_joined:
                x_hat.record_stream(main)
        parts = [out_z, latent]
        # out_lin is a RETURNED tensor only (eval_scores collects it, train() never reads it): the stacked fast path of
        # train.step_loss does not materialise the concatenation
        out_lin = None if stacked else torch.cat(parts, -1).detach()
        rparts = parts
        if self.isuseProb4Regr:
            if img_feat is None:
                img_feat = (x.view(B, self.rois, -1) * self.prob).reshape(B, -1)
            rparts = parts + [img_feat]                      # B rows: cat_linear reads it for both stacked passes
        # lin1 reads a column prefix of what lin1_regr reads: one operand-split launch serves both products, and the two products
        # (GEMM + split-K reduce each) run side by side on two streams
        ops.presplit_heads(rparts, [self.lin1.weight, self.lin1_regr.weight])
        aux = ops._aux_stream(x.device) if (_TWO_STREAMS and ops._AUX_STREAM) else None
        if aux is not None:
            aux.wait_stream(main)
            with torch.cuda.stream(aux):
                r = ops.cat_linear(rparts, self.lin1_regr.weight, self.lin1_regr.bias, relu=True)
            for t in rparts + ([ops._PRESPLIT["A"]] + list(ops._PRESPLIT["weights"].values()) if ops._PRESPLIT else []):
                t.record_stream(aux)
        linear_outf = ops.cat_linear(parts, self.lin1.weight, self.lin1.bias, relu=True)
        if aux is not None:
            main.wait_stream(aux)
            r.record_stream(main)
        else:
            r = ops.cat_linear(rparts, self.lin1_regr.weight, self.lin1_regr.bias, relu=True)
        # dropout, lin2 + log_softmax and lin2_regr of both heads in one launch (glue.cu)
        logp, our_reg = ops.output_heads(linear_outf, self._mask_of("lin1", linear_outf, 0.5), r, self._mask_of("lin1_regr", r, 0.3),
                                         self.lin2, self.lin2_regr)
        if side is not None and not go.decoder_joined:
            main.wait_stream(branch)             # x_hat (reconstruction loss) is needed by the caller, after the heads are queued
            x_hat.record_stream(main)
            go.decoder_joined = True
        if use_bank:
            bank.end_pass()
        outs = (logp, x_hat, out_z, out_lin, linear_outf, our_reg)
        if stacked:
            return outs                # rows [0, B) = plain pass, rows [B, 2B) = explain pass (out_lin: None, see above)
        return tuple(t[:B] for t in outs), tuple(t[B:] for t in outs)

    def _go_stream(self, device, which=0):
        sts = getattr(self, "_side_streams", None)
        if sts is None or sts[0].device != device:
            # the GO encoder chain (0) and the latent read-outs (3) are the long poles of the step: their CTAs are scheduled ahead of
            # the wide SGCN / Laplacian kernels they share the SMs with (stream priority is captured into the graph's kernel nodes)
            hi = os.environ.get("IGCN_GO_PRIORITY", "1") == "1"
            sts = [torch.cuda.Stream(device=device, priority=-1 if (hi and i in (0, 3)) else 0) for i in range(4)]
            self._side_streams = sts
        return sts[which]

    def supports_pair(self):
        return bool(self.isCrossAtten and not self.isImageOnly and not self.isSNPsOnly and not self.graph_pool)

    def __repr__(self):
        return self.__class__.__name__
