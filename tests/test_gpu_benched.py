"""Parity of the thing that is BENCHMARKED, at the benchmarked size: BASELINE configs[1] exactly as bench.py runs it -- B=256
graphs of 90 ROIs, `forward_pair` (plain + explain pass stacked), three streams, the whole step (zero_grad, forwards, losses,
backward, fused Adam) captured in one CUDA graph and replayed -- against the oracle's `train_step_loss` + torch Adam on the CPU,
in fp32 (what the reference would print) and fp64 (the truth).  Dropout is replayed from fixed masks on both sides.

BatchNorm batch statistics, the B x B consistency loss, the per-CTA partial-sum reductions and the split-K choices all depend on
the batch size, so the small golden fixtures do not cover this; the vectorised oracle runs B=256 in seconds.
Reference: kernel/train_eval_sgcn_img_snps.py:511-548, kernel/sgcn_img_snp.py:207-307.
"""
import numpy as np
import pytest
import torch

from oracle import igcn_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mask_shapes(B, pool):
    G, top = sum(pool), sum(pool[2:])
    shapes = dict(go_enc0=(B, G, 1), go_enc1=(B, G - pool[0], 1), go_B=(B, top), go_dec0=(B, G - pool[0], 1), go_dec1=(B, G, 1),
                  go_BD=(B, G), go_latent=(B, 32), lin1=(B, 64), lin1_regr=(B, 64))
    ps = dict(go_enc0=0.4, go_enc1=0.4, go_B=0.5, go_dec0=0.4, go_dec1=0.4, go_BD=0.5, go_latent=0.5, lin1=0.5, lin1_regr=0.3)
    return shapes, ps


def _oracle_run(P0, prep, c, w, lam, masks_p, masks_e, dtype, steps, patterns=None, flips=None):
    """patterns: per step ((enc, attn) of the plain pass, (enc, attn) of the explain pass) -- the ReLU active sets of the CUDA run,
    imposed on the oracle (oracle._relu); flips: list that receives (count, max |pre-activation| of a flipped element) per step."""
    P = {k: v.detach().clone().to(dtype).requires_grad_(True) if (v.is_floating_point() and "running" not in k) else v.clone()
         for k, v in P0.items()}
    b = {k: torch.from_numpy(v) for k, v in c.items()}
    for k in ("x", "edge_attr", "snps_feat", "clini_score", "tsne_fdim"):
        b[k] = b[k].to(dtype)
    leaves = [v for v in P.values() if v.requires_grad]
    opt = torch.optim.Adam(leaves, lr=1e-3)
    mp = {k: v.to(dtype) for k, v in masks_p.items()}
    me = {k: v.to(dtype) for k, v in masks_e.items()}
    losses, grads1 = [], None
    for s in range(steps):
        opt.zero_grad()
        b["x"] = b["x"].detach().requires_grad_(True)
        pat, probes = None, None
        if patterns is not None:
            pat = tuple(dict(enc=pp[0], attn=pp[1]) for pp in patterns[s])
            probes = ({}, {})
        loss, _, _ = O.train_step_loss(P, prep, b, w["L"], w["R"], lam, 0.01, True, mp, me, with_orth=False, patterns=pat, probes=probes)
        if flips is not None and probes is not None:
            cnt, worst = 0, 0.0
            for pr, pp in zip(probes, pat):
                for site in ("enc", "attn"):
                    f = pp[site] != (pr[site] > 0)
                    cnt += int(f.sum())
                    if bool(f.any()):
                        worst = max(worst, float(pr[site][f].abs().max() / pr[site].abs().max()))
            flips.append((cnt, worst))
        loss.backward()
        if s == 0:
            grads1 = {k: v.grad.detach().clone() for k, v in P.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}
        opt.step()
        losses.append(float(loss.detach()))
    return losses, grads1, {k: v.detach() for k, v in P.items()}


def _adam_params_close(got, p32, p64, g64, steps, lr):
    """Parameters after `steps` Adam updates.  Adam divides every coordinate by the root of its own second moment, so each live
    coordinate moves ~lr per step whatever its gradient's size, and a coordinate whose gradient g_i is small against the tensor
    inherits the RELATIVE error of g_i.  A gradient that meets the parity bar (|dg_i| <= RTOL (|g_i| + rms g), tests/helpers.py)
    therefore admits a displacement error of steps * lr * min(1, c * RTOL (|g_i| + rms g) / |g_i|) on coordinate i -- up to a full
    step where g_i vanishes (e.g. the attention's key bias, whose gradient is identically zero and whose update follows the sign of
    rounding noise, in the reference as much as here).  c = 4 covers the m / sqrt(v) ratio over the steps.  On top of that the
    usual rule A / rule B applies to the parameter values themselves.  Returns None or a failure message."""
    got, p32, p64, g = got.double(), p32.double(), p64.double(), g64.double().abs()
    rms_g = float(torch.sqrt((g * g).mean()))
    sens = H.RTOL * (g + rms_g) / g.clamp_min(1e-300)
    slack = steps * lr * torch.clamp(4.0 * sens, max=1.0)
    rms_p = float(torch.sqrt((p32 * p32).mean()))
    tol = slack + H.RTOL * (p32.abs() + rms_p)
    bad32 = (got - p32).abs() > tol
    if not bool(bad32.any()):
        return None
    # rule B on the offending coordinates: not further from the fp64 run than the fp32 reference is, by more than 2x
    eg, er = (got - p64).abs()[bad32], (p32 - p64).abs()[bad32]
    if bool((eg <= 2.0 * er + tol[bad32]).all()):
        return None
    i = int(((got - p32).abs() - tol).argmax())
    return "coordinate %d: |got - ref32| = %.3e > %.3e (|g_i| / rms g = %.2e)" % (
        i, float((got - p32).abs().flatten()[i]), float(tol.flatten()[i]), float(g.flatten()[i]) / max(rms_g, 1e-300))


def _relu_patterns(model, batch, B, dev):
    """ReLU active sets of the CUDA path for the current parameters: one eager no-grad forward_pair (same kernels as the captured
    step), BatchNorm buffers restored afterwards.  Returns ((enc, attn) plain, (enc, attn) explain) as CPU bool tensors."""
    from igcn_b200 import ops
    cap = {}
    orig = ops.cross_attention_average

    def spy(q, kv, mha):
        out = orig(q, kv, mha)
        cap["bx"], cap["oz"] = q.detach(), out.detach()
        return out

    saved = {k: v.detach().clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}
    ops.cross_attention_average = spy
    try:
        with torch.no_grad():
            model.forward_pair(batch, None, dev, stacked=True)
    finally:
        ops.cross_attention_average = orig
    model.load_state_dict(saved, strict=False)
    model._pe_cache = None
    torch.cuda.synchronize()
    enc = (cap["bx"] > 0).cpu()
    attn = ((2.0 * cap["oz"] - cap["bx"]) > 0).cpu()
    return (enc[:B], attn[:B]), (enc[B:], attn[B:])


@pytest.mark.parametrize("workload,B", [("config2", 256), ("config4", 24)])
def test_benched_graphed_step_vs_oracle(workload, B):
    """config2: the benchmarked configuration exactly (B=256, R=90).  config4: the same step at 264 ROIs (BASELINE configs[3]'s
    graph size) on a batch the CPU oracle finishes in seconds."""
    import bench
    from igcn_b200 import train as T
    from igcn_b200.data import Batch, SubjectSet
    w = dict(bench.WORKLOADS[workload], B=B)
    lam = list(bench.LAMBDA)
    dev = torch.device(DEV)
    model, sub, (adj, go_snps, pool_dim) = bench.build_problem(w, 0, dev)
    model = model.to(dev).train()
    P0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    shapes, ps = _mask_shapes(B, w["pool"])
    gen = torch.Generator().manual_seed(17)
    masks_p = {k: (torch.rand(s, generator=gen) >= ps[k]).float() / (1 - ps[k]) for k, s in shapes.items()}
    masks_e = {k: (torch.rand(s, generator=gen) >= ps[k]).float() / (1 - ps[k]) for k, s in shapes.items()}
    model.dropout_masks = {k: torch.cat([masks_p[k], masks_e[k]], 0).to(dev) for k in shapes}     # rows [0,B) plain, [B,2B) explain
    steps = 3
    # ---- the benchmarked path: FlatAdam + GraphedTrainStep (forward_pair, 3 streams, one CUDA graph) -------------------------
    opt = T.FlatAdam(model.parameters(), lr=1e-3)
    batch = Batch.collate(SubjectSet(sub), np.arange(B), dev)
    gs = T.GraphedTrainStep(model, opt, batch, lam, None, True)
    names = [n for n, _ in model.named_parameters()]
    losses, grads1, patterns = [], None, []
    for s in range(steps):
        patterns.append(_relu_patterns(model, batch, B, dev))
        losses.append(float(gs()))
        if s == 0:
            torch.cuda.synchronize()
            grads1 = {n: v.detach().cpu().clone() for n, v in zip(names, opt.grad_views)}
    torch.cuda.synchronize()
    final = {n: p.detach().cpu() for n, p in model.named_parameters()}
    # ---- the oracle, fp32 and fp64 ---------------------------------------------------------------------------------------------
    prep = O.go_index_prep(adj.T, go_snps, w["pool"])
    c = O.collate(sub, np.arange(B))
    flips = []
    l32, g32, p32 = _oracle_run(P0, prep, c, w, lam, masks_p, masks_e, torch.float32, steps, patterns)
    l64, g64, p64 = _oracle_run(P0, prep, c, w, lam, masks_p, masks_e, torch.float64, steps, patterns, flips)
    tag = "%s B=%d graphed: " % (workload, B)
    # the CUDA run's ReLU active sets (encoder layers, attention output: 2.9 M elements at B=256) were imposed on the oracle; the
    # elements where the fp64 oracle would have decided otherwise must be few and within rounding distance of zero
    n_el = 2 * 2 * B * w["R"] * w["L"] * w["H"]
    for cnt, worst in flips:
        assert cnt <= max(3, int(1e-5 * n_el)) and worst <= 2e-5, (cnt, worst)
    H.PARITY_LOG.append(dict(what=tag + "ReLU sign flips vs fp64 per step %s" % ([f[0] for f in flips],), rule="A", err32=float(sum(f[0] for f in flips))))
    with H.Collector() as col:
        col.parity(np.asarray(losses), np.asarray(l32), np.asarray(l64), what=tag + "loss trajectory")
        assert len(g64) >= 40, len(g64)
        for k, g in g64.items():
            col.parity(grads1[k], g32[k], g, what=tag + "step-1 grad " + k)
        for k in g64:                   # every parameter that receives a gradient, after 3 fused Adam updates
            msg = _adam_params_close(final[k], p32[k], p64[k], g64[k], steps, 1e-3)
            if msg:
                col.failures.append(tag + "after %d Adam steps: %s: %s" % (steps, k, msg))
    for k, v in final.items():          # ... and the others did not move
        if k not in g64:
            assert torch.equal(v, P0[k]), k


def test_lr_decay_is_honoured_by_eager_and_graphed_steps():
    """The reference halves param_groups[0]['lr'] in place (train_eval_sgcn_img_snps.py:169-171); both step forms must follow."""
    import bench
    from igcn_b200 import train as T
    from igcn_b200.data import Batch, SubjectSet
    w = dict(bench.WORKLOADS["config2"], B=8)
    dev = torch.device(DEV)
    model, sub, _ = bench.build_problem(w, 0, dev)
    model = model.to(dev).train()
    model.dropout_masks = {k: torch.ones(1, device=dev) for k in O.MODEL_MASK_NAMES}
    opt = T.FlatAdam(model.parameters(), lr=1e-3)
    batch = Batch.collate(SubjectSet(sub), np.arange(8), dev)
    for graphed in (False, True):
        step = T.GraphedTrainStep(model, opt, batch, bench.LAMBDA, None, True) if graphed else \
            (lambda: T.train_step(model, batch, opt, bench.LAMBDA, None, True))
        moved = []
        for lr in (1e-3, 0.0):
            opt.param_groups[0]["lr"] = lr
            before = opt.flat_param.clone()
            step()
            torch.cuda.synchronize()
            moved.append(float((opt.flat_param - before).abs().max()))
        assert moved[0] > 0.0 and moved[1] == 0.0, (graphed, moved)
