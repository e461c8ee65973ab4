"""The training step of the hot path (reference: kernel/train_eval_sgcn_img_snps.py:511-548) and its
data-parallel form: one process per GPU, graphs sharded across ranks, ONE flat fp32 gradient all-reduce per step.
"""
from __future__ import annotations

import os
import types

import torch
import torch.distributed as dist
import torch.nn.functional as F

# sgcn_hyperparameters.py:18-23
hp = types.SimpleNamespace(lamda_x_l1=0.1, lamda_e_l1=0.1, lamda_x_ent=0.1, lamda_e_ent=0.1, lamda_mi=1, lamda_ce=1)

DEFAULT_LAMBDA = [0.0, 1.0, 0.5, 0.0000015, 0.1, 0.0]      # main.py:73-78 -> main.py:204


def step_loss(model, data, lambda_loss=None, isSoftSimilarity=True, temperature=None, num_cluster=2, hyper=hp, pair=True,
              return_logp=False):
    """The scalar that train() back-propagates, term by term as in train_eval_sgcn_img_snps.py:521-544 (eval_loss evaluates the same
    expression, :564-598).  return_logp: also return the plain pass's log-probabilities (B, C)."""
    lam = DEFAULT_LAMBDA if lambda_loss is None else lambda_loss
    dev = data.x.device
    y = data.y.view(-1)
    cs = data.clini_score.view(-1)
    snps = data.snps_feat
    if pair and isSoftSimilarity and getattr(model, "supports_pair", lambda: False)() and data.x.is_cuda:
        # both passes in one sweep on 2B stacked samples (identical results, half the launches; SGCN_GCN_IMGSNP.forward_pair);
        # every loss term of train() is the mean of its plain and explain values, so it is evaluated on the stacked tensors
        # directly and no slicing enters the autograd graph
        out2, snps_hat2, out_feat2, _, _, our_reg2 = model.forward_pair(data, temperature, dev, stacked=True, consist=True,
                                                                        mask_loss_hp=hyper)
        B = data.snps_feat.shape[0]
        from . import ops
        c, model._lp_cache = getattr(model, "_lp_cache", None), None
        if c is not None and c[0] is hyper:              # already queued beside the attention by forward_pair
            lp = c[1]
            cur = torch.cuda.current_stream(dev)
            cur.wait_stream(c[2])
            lp.record_stream(cur)
        else:
            lp = model.loss_probability(data.x, data.edge_index, data.edge_attr, hyper)
        quad = model.consist_loss_pair(out_feat2, data.tsne_fdim)
        # lam1 * (mse + mse_p)/2 + lam2 * loss_prob + lam3 * (recon + recon_p)/2 + lam4 * (cluster + cluster_p)/2 as one launch
        loss_reg = ops.step_loss_pair(our_reg2, cs, snps_hat2, snps, lp, quad, lam[1], lam[3] / 2, lam[2], lam[4] / 2)
        loss_prob = recon = cluster = 0.0
        out, out_p, out_feat = out2[:B], out2[B:], out_feat2[:B]
    else:
        if pair and hasattr(model, "forward_pair"):
            (out, snps_hat, out_feat, out_lin, _, our_reg), (out_p, snps_hat_p, out_feat_p, out_lin_p, _, our_reg_p) = \
                model.forward_pair(data, temperature, dev)
        else:
            out, snps_hat, out_feat, out_lin, _, our_reg = model(data, temperature, dev)
            out_p, snps_hat_p, out_feat_p, out_lin_p, _, our_reg_p = model(data, temperature, dev, isExplain=True)
        loss_reg = lam[1] * (F.mse_loss(our_reg.view(-1), cs) + F.mse_loss(our_reg_p.view(-1), cs)) / 2
        loss_prob = lam[2] * model.loss_probability(data.x, data.edge_index, data.edge_attr, hyper)
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
    loss = None
    for term in (loss_reg, loss_prob, recon, cluster, orth):
        if torch.is_tensor(term):                     # python zeros stand for terms that are absent or already folded in
            loss = term if loss is None else loss + term
    if loss is None:
        loss = torch.zeros((), device=dev)
    if lam[0] != 0:
        loss = loss + hyper.lamda_ce * lam[0] * F.nll_loss(out, y) + hyper.lamda_mi * lam[0] * F.nll_loss(out_p, y)
    return (loss, out) if return_logp else loss


def evaluate(model, loader, lambda_loss=None, isSoftSimilarity=True, temperature=None, hyper=hp):
    """eval_loss and eval_acc of the reference (kernel/train_eval_sgcn_img_snps.py:551-600) in ONE sweep over `loader`.

    The reference runs them as separate loops (plus eval_scores): five eval-mode forwards per batch and a `.cpu().item()` per batch.
    Here a batch is one stacked plain + explain inference pass (forward_pair under model.eval(): running-statistics BatchNorm as a
    fused affine kernel, no dropout), the weighted loss and the number of correct predictions accumulate ON THE DEVICE, and the
    host reads two scalars at the end.  Returns (mean loss per graph, accuracy).  Leaves the model in eval mode, as the reference."""
    model.eval()
    dev = next(model.parameters()).device
    loss_sum = torch.zeros((), dtype=torch.float64, device=dev)
    correct = torch.zeros((), dtype=torch.int64, device=dev)
    n = 0
    with torch.no_grad():
        for data in loader:
            data = data.to(dev)
            loss, logp = step_loss(model, data, lambda_loss, isSoftSimilarity, temperature, hyper=hyper, return_logp=True)
            b = int(data.num_graphs)
            loss_sum += loss.double() * b
            correct += (logp.argmax(1) == data.y.view(-1)).sum()
            n += b
            model._pe_cache = None
            model._w_cache = None
    if n == 0:
        return 0.0, 0.0
    host = torch.stack([loss_sum, correct.double()]).cpu()
    return float(host[0]) / n, float(host[1]) / n


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


class FlatAdam(object):
    """Adam over ONE flat fp32 parameter buffer, one fused kernel per step (igcn_adam_step).

    Every parameter is re-pointed to a view of `flat_param` (names, shapes and state_dict are unchanged); gradients are
    gathered into `flat_grad` -- the buffer the data-parallel all-reduce runs on -- by one multi-tensor copy after
    backward, so autograd never pays a per-parameter `grad += g` kernel for the first use of a parameter.
    Semantics = torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay=0) as the reference uses it
    (kernel/train_eval_sgcn_img_snps.py:108); parameters that receive no gradient keep zero moments and do not move.
    `param_groups[0]['lr']` is honoured (the reference decays it in place, :169-171)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, process_group=None):
        self.params = [p for p in params]
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam runs on CUDA parameters only (no CPU fallback)")
        sizes = [p.numel() for p in self.params]
        self.offsets, off = [], 0
        for n in sizes:
            self.offsets.append(off)
            off += (n + 3) // 4 * 4                    # keep every parameter 16-byte aligned inside the flat buffers
        self.n = off
        f = lambda: torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq = f(), f(), f(), f()
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                v = self.flat_param[o:o + p.numel()].view_as(p)
                v.copy_(p.data)
                p.data = v
        self.grad_views = [self.flat_grad[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)]
        self.step_t = torch.zeros(1, dtype=torch.float32, device=dev)
        self.lr_t = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.param_groups = [dict(params=self.params, lr=float(lr), betas=betas, eps=eps)]
        self._lr_seen = float(lr)
        self.group = process_group
        self._peer = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1 \
                and os.environ.get("IGCN_DP", "peer") != "nccl":
            self._enable_peer_allreduce()

    def _enable_peer_allreduce(self):
        """Move the flat gradient buffer into symmetric (peer-mapped) memory so the gradient all-reduce and the Adam update run
        as ONE kernel over NVLink (igcn_dp_allreduce_adam).  Falls back to ncclAllReduce + igcn_adam_step when the peer mapping
        cannot be established (e.g. no P2P access between the ranks' devices)."""
        import ctypes
        import sys
        from . import _lib
        try:
            import torch.distributed._symmetric_memory as symm_mem
            group = self.group if self.group is not None else dist.group.WORLD
            dev = self.flat_param.device
            buf = symm_mem.empty(self.n, dtype=torch.float32, device=dev)
            hdl = symm_mem.rendezvous(buf, group)
            world, rank = int(hdl.world_size), int(hdl.rank)
            pad = int(hdl.signal_pad_size)
            if _lib.lib().igcn_dp_adam_blocks(self.n, world, pad) < 1:
                raise RuntimeError("signal pad of %d bytes too small for %d ranks" % (pad, world))
            buf.zero_()
            self.flat_grad = buf
            self.grad_views = [buf[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)]
            gp = (ctypes.c_int64 * world)(*[int(a) for a in hdl.buffer_ptrs])
            sp = (ctypes.c_int64 * world)(*[int(a) for a in hdl.signal_pad_ptrs])
            self._peer = dict(handle=hdl, grad_ptrs=gp, signal_ptrs=sp, world=world, rank=rank, pad=pad,
                              timeout_ms=int(float(os.environ.get("IGCN_DP_TIMEOUT_S", "120")) * 1000),
                              error=torch.zeros(1, dtype=torch.int32, device=dev))
            torch.cuda.synchronize(dev)
        except Exception as e:                                     # noqa: BLE001 -- any failure means "use NCCL"
            self._peer = None
            sys.stderr.write("igcn_b200.FlatAdam: peer-memory all-reduce unavailable (%s: %s); using ncclAllReduce\n" % (type(e).__name__, e))
        # all ranks must take the same path (a rank that fell back would never answer the others' flags)
        ok = torch.tensor([1 if self._peer is not None else 0], dtype=torch.int32, device=self.flat_param.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0 and self._peer is not None:
            self._peer = None
            self.flat_grad = torch.zeros(self.n, dtype=torch.float32, device=self.flat_param.device)
            self.grad_views = [self.flat_grad[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)]

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def rendezvous(self):
        """Host barrier for all ranks: call before the first fused step and after long rank-local host phases (evaluation,
        checkpointing) so the device-side flag waits of igcn_dp_allreduce_adam start within their timeout of each other."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            torch.cuda.synchronize(self.flat_param.device)
            dist.barrier(group=self.group)

    def check_dp_error(self):
        """Raises if a device-side wait of the fused all-reduce timed out (synchronises)."""
        if self._peer is not None:
            code = int(self._peer["error"].item())
            if code:
                raise RuntimeError("igcn_dp_allreduce_adam: %s rank %d timed out after %d ms; parameters are no longer in sync"
                                   % ("signal to" if (code >> 8) == 1 else "wait for", (code & 255) - 1, self._peer["timeout_ms"]))

    def sync_lr(self):
        """Push param_groups[0]['lr'] (the reference decays it in place, train_eval_sgcn_img_snps.py:169-171) into the device
        scalar the kernels read.  step() calls it whenever the stream is not capturing and GraphedTrainStep calls it before every
        replay, so a changed learning rate is honoured by the eager and by the graphed step (the graph reads the device scalar)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_seen:
            self.lr_t.fill_(lr)
            self._lr_seen = lr

    def gather_grads(self, count_step=False):
        """p.grad of every parameter -> its slot of flat_grad, one launch (igcn_gather_flat); a parameter without a gradient gets
        zeros.  Gradients that are not plain contiguous f32 tensors (none in this package's models) take a torch copy.
        count_step: the same launch increments the device step counter the update reads."""
        from . import _lib
        import ctypes
        n = len(self.params)
        src, odd = (ctypes.c_int64 * n)(), []
        for i, p in enumerate(self.params):
            g = p.grad
            if g is None:
                src[i] = 0
            elif g.dtype == torch.float32 and g.is_contiguous() and g.device == self.flat_grad.device and not g.is_sparse:
                src[i] = g.data_ptr()
            else:
                src[i] = 0
                odd.append((self.grad_views[i], g))
        if getattr(self, "_gather_tab", None) is None:
            self._gather_tab = ((ctypes.c_int64 * n)(*self.offsets), (ctypes.c_int64 * n)(*[p.numel() for p in self.params]))
        off, sizes = self._gather_tab
        with torch.cuda.device(self.flat_grad.device):
            _lib.call("igcn_gather_flat", ctypes.addressof(src), ctypes.addressof(off), ctypes.addressof(sizes), n,
                      _lib.ptr(self.flat_grad), self.n, _lib.ptr(self.step_t) if count_step else None, _lib.stream(),
                      tag="gather_flat", nbytes=8 * self.n)
            _lib.launch_count += int(_lib.lib().igcn_gather_flat_launches(n)) - 1
        for v, g in odd:
            v.copy_(g)

    def step(self):
        from . import _lib
        import ctypes
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        self.gather_grads(count_step=True)
        if self._peer is not None:
            # gradient all-reduce (sum in rank order, / world) + Adam in one kernel over peer memory
            g, pr = self.param_groups[0], self._peer
            with torch.cuda.device(self.flat_param.device):
                _lib.call("igcn_dp_allreduce_adam", ctypes.addressof(pr["grad_ptrs"]), ctypes.addressof(pr["signal_ptrs"]), pr["rank"], pr["world"],
                          pr["pad"], _lib.ptr(self.flat_param), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), _lib.ptr(self.step_t),
                          _lib.ptr(self.lr_t), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), self.n, pr["timeout_ms"],
                          _lib.ptr(pr["error"]), _lib.stream(),
                          tag="dp_allreduce_adam", nbytes=4 * self.n * (pr["world"] + 6))
            return
        scale = 1.0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1 \
                and not os.environ.get("IGCN_DIAG_NO_ALLREDUCE"):       # diagnostic switch: isolates the collective's cost
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            scale = 1.0 / dist.get_world_size(self.group)
        g = self.param_groups[0]
        with torch.cuda.device(self.flat_param.device):
            _lib.call("igcn_adam_step", _lib.ptr(self.flat_param), _lib.ptr(self.flat_grad), _lib.ptr(self.exp_avg),
                      _lib.ptr(self.exp_avg_sq), _lib.ptr(self.step_t), _lib.ptr(self.lr_t), float(g["betas"][0]),
                      float(g["betas"][1]), float(g["eps"]), scale, self.n, _lib.stream())


_SEEDS = {}


def _grad_seed(loss):
    key = (loss.device, loss.dtype, tuple(loss.shape))
    t = _SEEDS.get(key)
    if t is None:
        if loss.is_cuda and torch.cuda.is_current_stream_capturing():
            return None                   # never allocate the persistent seed inside a capture; the warm-up steps create it
        t = torch.ones_like(loss)
        _SEEDS[key] = t
    return t


def train_step(model, data, optimizer=None, lambda_loss=None, flat: FlatGradAllReduce = None, isSoftSimilarity=True):
    """zero_grad -> 2 forwards + losses -> backward -> (all-reduce) -> optimizer.step. Returns the detached loss."""
    if flat is not None:
        flat.zero()
    elif optimizer is not None:
        optimizer.zero_grad()
    if data.x.grad is not None:
        data.x.grad = None
    loss = step_loss(model, data, lambda_loss, isSoftSimilarity)
    from . import ops
    ops.defer_weight_grad_joins()         # weight-gradient products on auxiliary streams are joined once, here, before the optimizer
    try:
        loss.backward(_grad_seed(loss))   # a cached ones tensor: torch would launch a fill for the seed, on the step's critical path
    finally:
        ops.join_weight_grads()
    if getattr(model, "_pe_cache", None) is not None:
        model._pe_cache = None            # drop the last reference to this step's autograd graph
    if getattr(model, "_w_cache", None) is not None:
        model._w_cache = None             # the similarity matrix belongs to this batch only
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
        if isinstance(optimizer, FlatAdam):
            optimizer.sync_lr()
            optimizer.rendezvous()                # ranks may arrive here seconds apart (lazy loads, data set-up)
        self.model, self.opt, self.batch, self.flat = model, optimizer, static_batch, flat
        self.lambda_loss, self.soft = lambda_loss, isSoftSimilarity
        dev = static_batch.x.device
        model._pe_cache = None
        # The warm-up steps below are real training steps (they must be: lazy initialisation, the dropout-mask plan and the
        # allocator's pools have to be in their steady state before capture).  Their effect is undone: parameters, BatchNorm
        # buffers, Adam moments / step count and the dropout counter are snapshotted here and restored IN PLACE after capture,
        # so building the graph leaves the training trajectory exactly where it was.
        snap = self._snapshot()
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
        self._restore(snap)
        torch.cuda.synchronize(dev)

    def _mask_bank(self):
        go = getattr(self.model, "go_network", None)
        return getattr(go, "mask_bank", None)

    def _snapshot(self):
        import copy
        snap = dict(model={k: v.detach().clone() for k, v in self.model.state_dict().items()})
        if isinstance(self.opt, FlatAdam):
            o = self.opt
            snap["adam"] = (o.exp_avg.clone(), o.exp_avg_sq.clone(), o.step_t.clone())
        elif self.opt is not None:
            snap["opt"] = copy.deepcopy(self.opt.state_dict())
        bank = self._mask_bank()
        snap["counter"] = None if bank is None or bank.counter is None else bank.counter.clone()
        return snap

    def _restore(self, snap):
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(snap["model"][k])                      # in place: the captured graph holds these addresses
            if "adam" in snap:
                o = self.opt
                o.exp_avg.copy_(snap["adam"][0])
                o.exp_avg_sq.copy_(snap["adam"][1])
                o.step_t.copy_(snap["adam"][2])
            elif "opt" in snap:
                cur = self.opt.state_dict()
                old = snap["opt"]
                for pid, st in cur["state"].items():           # in place as well (capturable optimizers keep device state)
                    for name, val in st.items():
                        if torch.is_tensor(val):
                            if pid in old["state"] and name in old["state"][pid]:
                                val.copy_(old["state"][pid][name])
                            else:
                                val.zero_()
            bank = self._mask_bank()
            if bank is not None and bank.counter is not None:
                if snap["counter"] is not None:
                    bank.counter.copy_(snap["counter"])
                else:
                    bank.counter.zero_()

    def __call__(self):
        if isinstance(self.opt, FlatAdam):
            self.opt.sync_lr()                    # lr_t.fill_ outside the graph; the replay reads the device scalar
        self.graph.replay()
        return self.loss
