"""The training step of the hot path (reference: kernel/train_eval_sgcn_img_snps.py:511-548) and its
data-parallel form: one process per GPU, graphs sharded across ranks, ONE flat fp32 gradient all-reduce per step.
"""
from __future__ import annotations

import types

import torch
import torch.distributed as dist
import torch.nn.functional as F

# sgcn_hyperparameters.py:18-23
hp = types.SimpleNamespace(lamda_x_l1=0.1, lamda_e_l1=0.1, lamda_x_ent=0.1, lamda_e_ent=0.1, lamda_mi=1, lamda_ce=1)

DEFAULT_LAMBDA = [0.0, 1.0, 0.5, 0.0000015, 0.1, 0.0]      # main.py:73-78 -> main.py:204


def step_loss(model, data, lambda_loss=None, isSoftSimilarity=True, temperature=None, num_cluster=2, hyper=hp):
    """The scalar that train() back-propagates, term by term as in train_eval_sgcn_img_snps.py:521-544."""
    lam = DEFAULT_LAMBDA if lambda_loss is None else lambda_loss
    dev = data.x.device
    y = data.y.view(-1)
    out, snps_hat, out_feat, out_lin, _, our_reg = model(data, temperature, dev)
    out_p, snps_hat_p, out_feat_p, out_lin_p, _, our_reg_p = model(data, temperature, dev, isExplain=True)
    cs = data.clini_score.view(-1)
    loss_reg = lam[1] * (F.mse_loss(our_reg.view(-1), cs) + F.mse_loss(our_reg_p.view(-1), cs)) / 2
    loss_prob = lam[2] * model.loss_probability(data.x, data.edge_index, data.edge_attr, hyper)
    snps = data.snps_feat
    recon = lam[3] * (((snps_hat - snps) ** 2).sum() + ((snps_hat_p - snps) ** 2).sum()) / 2
    cluster = 0
    if isSoftSimilarity:
        cluster = lam[4] * (model.consist_loss(out_feat, data.tsne_fdim) + model.consist_loss(out_feat_p, data.tsne_fdim)) / 2
    else:
        for c in range(num_cluster):
            m = data.clust_y == c
            cluster = cluster + lam[4] * (model.consist_loss(out_feat[m]) + model.consist_loss(out_feat_p[m])) / 2
    # the reference evaluates OrthogonalConstraint even when its weight is 0 (train_eval...:538); skipping a
    # zero-weighted term changes neither the loss nor any gradient
    orth = lam[5] * model.OrthogonalConstraint(out_feat) if lam[5] != 0 else 0.0
    loss = loss_reg + loss_prob + recon + cluster + orth
    if lam[0] != 0:
        loss = loss + hyper.lamda_ce * lam[0] * F.nll_loss(out, y) + hyper.lamda_mi * lam[0] * F.nll_loss(out_p, y)
    return loss


class FlatGradAllReduce(object):
    """Data-parallel plumbing: all parameter gradients live as views of ONE flat fp32 buffer; a step does a
    single all-reduce(SUM) on it and scales by 1/world.  Parameters that never get a
    gradient (edge_prob, batch_norm*, go_network.classification.*) keep zeroed slots."""

    def __init__(self, model, process_group=None):
        self.params = [p for p in model.parameters()]
        self.group = process_group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()
        off = 0
        for p in self.params:          # re-attach in case an optimizer/zero_grad(set_to_none) dropped the views
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def reduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / dist.get_world_size(self.group))


def train_step(model, data, optimizer=None, lambda_loss=None, flat: FlatGradAllReduce = None, isSoftSimilarity=True):
    """zero_grad -> 2 forwards + losses -> backward -> (all-reduce) -> optimizer.step. Returns the detached loss."""
    if flat is not None:
        flat.zero()
    elif optimizer is not None:
        optimizer.zero_grad()
    loss = step_loss(model, data, lambda_loss, isSoftSimilarity)
    loss.backward()
    if getattr(model, "_pe_cache", None) is not None:
        model._pe_cache = None            # drop the last reference to this step's autograd graph
    if flat is not None:
        flat.reduce()
    if optimizer is not None:
        optimizer.step()
    return loss.detach()


class GraphedTrainStep(object):
    """The whole training step captured ONCE in a CUDA graph and replayed per batch.

    Every launch of the step (the igcn kernels through the C ABI, torch's small ops, the NCCL all-reduce and the
    capturable Adam update) is static for a fixed (B, E): after removing the reference's host syncs
    (`x.min().item()`, boolean-mask indexing in gcn_norm, `num_graphs`) nothing in the step depends on device data,
    so one graph launch replaces several hundred kernel launches.  New batches are collated straight into the
    graph's static input buffers (Batch.collate(out=...)).
    """

    def __init__(self, model, optimizer, static_batch, lambda_loss=None, flat: FlatGradAllReduce = None,
                 isSoftSimilarity=True, warmup=3):
        self.model, self.opt, self.batch, self.flat = model, optimizer, static_batch, flat
        self.lambda_loss, self.soft = lambda_loss, isSoftSimilarity
        dev = static_batch.x.device
        model._pe_cache = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                train_step(model, static_batch, optimizer, lambda_loss, flat, isSoftSimilarity)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        static_batch.x.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = train_step(model, static_batch, optimizer, lambda_loss, flat, isSoftSimilarity)

    def __call__(self):
        self.graph.replay()
        return self.loss
