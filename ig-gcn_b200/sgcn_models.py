"""Image-only SGCN model family on the fused kernels (reference: kernel/sgcn.py): `SGCN_GCN` (GCNConv stack) and
`SGCN_GAT` (GATConv(edge_dim=1) stack), same constructor / forward(data, isExplain) / cal_probability /
loss_probability surface and parameter names.  `rois` is honoured in lin1 (the reference hard-codes 90, sgcn.py:167,285)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import init
from torch.nn.parameter import Parameter

from . import ops
from .img_snp_model import MaskedEncoderMixin, _l1_entropy
from .pyg import GATConv, GCNConv          # the PyG-signature operators (parameter containers here: the models fuse the stack)


class _SGCNBase(nn.Module, MaskedEncoderMixin):
    def _init_masks(self):
        self.prob = Parameter(torch.zeros((self.rois, self.prob_dim)))
        self.prob_bias = Parameter(torch.empty((self.prob_dim * 2, 1)))
        init.kaiming_uniform_(self.prob_bias, a=math.sqrt(5))
        self.edge_prob = Parameter(torch.empty((self.rois, self.rois)))
        init.kaiming_uniform_(self.prob, a=math.sqrt(5))
        init.kaiming_uniform_(self.edge_prob, a=math.sqrt(5))

    def activations_hook(self, grad):
        self.final_conv_grads = grad

    def loss_probability(self, x, edge_index, edge_weight, hp, eps=1e-6):
        """kernel/sgcn.py:329-351: per-row L1 / N for the node mask, no SNP term."""
        edge_prob = self._edge_prob(x, edge_index, edge_weight)
        x_prob = torch.sigmoid(self.prob)
        f_sum = x_prob.abs().sum(-1).sum() / x_prob.shape[0]
        _, f_ent = _l1_entropy(x_prob, eps)
        e_sum, e_ent = _l1_entropy(edge_prob, eps)
        return hp.lamda_x_l1 * f_sum + hp.lamda_e_l1 * e_sum + hp.lamda_x_ent * f_ent + hp.lamda_e_ent * e_ent

    def _head(self, z):
        x = F.relu(self.lin1(z))
        x = F.dropout(x, p=0.5, training=self.training)
        return F.log_softmax(self.lin2(x), dim=-1)

    def __repr__(self):
        return self.__class__.__name__


class SGCN_GCN(_SGCNBase):
    def __init__(self, dataset, num_layers, hidden, *args, hidden_linear=64, rois=90, H_0=3, num_features=3, num_classes=2, **kwargs):
        super().__init__()
        self.input = self.final_conv_acts = self.final_conv_grads = None
        self.rois, self.prob_dim = rois, H_0
        self.conv1 = GCNConv(num_features, hidden)
        self.convs = nn.ModuleList([GCNConv(hidden, hidden) for _ in range(num_layers - 1)])
        self.lin1 = nn.Linear(rois * num_layers * hidden, hidden_linear)
        self.lin2 = nn.Linear(hidden_linear, num_classes)
        self._init_masks()

    def forward(self, data, isExplain=False):
        x = data.x
        if not x.requires_grad and x.is_leaf:
            x.requires_grad = True
        self.input = x
        csr = self._csr_for(data, self.rois)
        Ws, bs = self._conv_params()
        if isExplain:
            out, p_e = ops.sgcn_encoder(x, csr, Ws, bs, self.prob, self.prob_bias, want_pe=True)
            self._pe_cache = ((x.data_ptr(), data.edge_index.data_ptr(), self.prob._version, self.prob_bias._version), p_e,
                              torch.is_grad_enabled())
        else:
            out, _ = ops.sgcn_encoder(x, csr, Ws, bs)
        return self._head(out.view(out.shape[0], -1))


class SGCN_GAT(_SGCNBase):
    def __init__(self, dataset, num_layers, hidden, *args, hidden_linear=64, rois=90, H_0=3, **kwargs):
        super().__init__()
        self.input = self.final_conv_acts = self.final_conv_grads = None
        self.rois, self.prob_dim = rois, H_0
        self.conv1 = GATConv(dataset.num_features, hidden, edge_dim=1)
        self.convs = nn.ModuleList([GATConv(hidden, hidden, edge_dim=1) for _ in range(num_layers - 1)])
        self.lin1 = nn.Linear(rois * num_layers * hidden, hidden_linear)
        self.lin2 = nn.Linear(hidden_linear, dataset.num_classes)
        self._init_masks()

    def forward(self, data, isExplain=False):
        x, edge_index, edge_weight = data.x, data.edge_index, data.edge_attr
        if not x.requires_grad and x.is_leaf:
            x.requires_grad = True
        self.input = x
        csr = self._csr_for(data, self.rois)
        if isExplain:
            N, D = x.shape
            h = (x.view(N // self.rois, self.rois, D) * self.prob).reshape(N, D)
            p_e = ops.edge_mask(x, csr, self.prob, self.prob_bias)                       # CSR-slot order
            self._pe_cache = ((x.data_ptr(), edge_index.data_ptr(), self.prob._version, self.prob_bias._version), p_e,
                              torch.is_grad_enabled())
            ea = edge_weight.index_select(0, csr.csr_perm.long()) * p_e                  # masked weights per CSR slot
        else:
            h, ea = x, edge_weight.index_select(0, csr.csr_perm.long())
        xs = []
        for conv in [self.conv1] + list(self.convs):
            h = F.relu(ops.gat_conv(h, csr, ea, conv.lin_src.weight, conv.att_src, conv.att_dst, conv.lin_edge.weight,
                                    conv.att_edge, conv.bias, conv.negative_slope))
            xs.append(h)
        z = torch.cat(xs, dim=1)
        return self._head(z.view(csr.B, -1))
