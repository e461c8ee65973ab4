"""Data-parallel correctness on 2 GPUs (skipped on a single-GPU box): the fused peer-memory all-reduce + Adam kernel against
ncclAllReduce + igcn_adam_step on the same buffers, and a 2-rank training step against a single-process run that applies the
documented data-parallel rule (per-rank BatchNorm statistics and per-rank consistency loss, gradients averaged over the ranks;
DESIGN.md section 6).  Reference: one optimizer.step() per batch, kernel/train_eval_sgcn_img_snps.py:511-548,
made data parallel over graphs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import bench
    from igcn_b200 import _lib, train as T
    from igcn_b200.data import Batch, SubjectSet
    from oracle import igcn_oracle as O
    res = {}
    # ---- (1) fused all-reduce + Adam == ncclAllReduce + igcn_adam_step on the same gradients ------------------------------------
    B = 16
    w = dict(bench.WORKLOADS["config2"], B=B)
    model, _, _ = bench.build_problem(w, rank, dev)                      # identical replicas (seed 0 inside)
    model = model.to(dev).train()
    opt = T.FlatAdam(model.parameters(), lr=1e-3)
    res["peer_path"] = opt._peer is not None
    g = torch.Generator().manual_seed(100 + rank)
    grad = torch.randn(opt.n, generator=g).to(dev)
    ref_g = grad.clone()
    dist.all_reduce(ref_g, op=dist.ReduceOp.SUM)
    p_ref, m_ref, v_ref = opt.flat_param.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone()
    step_ref, lr_t = torch.ones(1, device=dev), torch.full((1,), 1e-3, device=dev)
    with torch.cuda.device(dev):
        _lib.call("igcn_adam_step", _lib.ptr(p_ref), _lib.ptr(ref_g), _lib.ptr(m_ref), _lib.ptr(v_ref), _lib.ptr(step_ref), _lib.ptr(lr_t),
                  0.9, 0.999, 1e-8, 1.0 / world, opt.n, _lib.stream())
    for p in opt.params:
        p.grad = None
    opt.flat_grad.copy_(grad)
    # the flat buffer already holds this rank's gradient; the gather launch also counts the step
    opt.gather_grads = lambda count_step=False: (opt.step_t.add_(1.0) if count_step else None)
    opt.step()
    torch.cuda.synchronize()
    opt.check_dp_error()
    res["allreduce_adam_max_abs_diff"] = float((opt.flat_param - p_ref).abs().max())
    res["allreduce_adam_moment_diff"] = float((opt.exp_avg - m_ref).abs().max())
    # ---- (2) a 2-rank training step == the documented rule evaluated in one process ----------------------------------------------
    del opt
    model, sub_r, _ = bench.build_problem(w, rank, dev)                  # rank r holds subjects [r*B, (r+1)*B)
    model = model.to(dev).train()
    shapes = dict(go_enc0=(2 * B, 54, 1), go_enc1=(2 * B, 34, 1), go_B=(2 * B, 19), go_dec0=(2 * B, 34, 1), go_dec1=(2 * B, 54, 1),
                  go_BD=(2 * B, 54), go_latent=(2 * B, 32), lin1=(2 * B, 64), lin1_regr=(2 * B, 64))
    ps = dict(go_enc0=0.4, go_enc1=0.4, go_B=0.5, go_dec0=0.4, go_dec1=0.4, go_BD=0.5, go_latent=0.5, lin1=0.5, lin1_regr=0.3)

    def masks_of(r):
        gg = torch.Generator().manual_seed(7 + r)
        return {k: ((torch.rand(s, generator=gg) >= ps[k]).float() / (1 - ps[k])).to(dev) for k, s in shapes.items()}

    model.dropout_masks = masks_of(rank)
    opt = T.FlatAdam(model.parameters(), lr=1e-3)
    batch = Batch.collate(SubjectSet(sub_r), np.arange(B), dev)
    losses = [float(T.train_step(model, batch, opt, bench.LAMBDA)) for _ in range(2)]
    torch.cuda.synchronize()
    opt.check_dp_error()
    res["dp_losses"] = losses
    dp_params = opt.flat_param.clone()
    if rank == 0:
        # single process: both shards, each with ITS masks and ITS BatchNorm statistics / consistency loss, gradients averaged
        from igcn_b200 import synthetic as syn
        ref_model, _, _ = bench.build_problem(w, 0, dev)
        ref_model = ref_model.to(dev).train()
        subs = [syn.make_subjects(B, rois=w["R"], n_snps=w["S"], seed=1234, first_id=r * B, num_classes=w["num_classes"], num_regr=w["num_regr"])
                for r in range(world)]
        batches = [Batch.collate(SubjectSet(s), np.arange(B), dev) for s in subs]
        params = [p for p in ref_model.parameters()]
        flat_p = torch.cat([p.detach().reshape(-1) for p in params]).double()
        m1, v1 = torch.zeros_like(flat_p), torch.zeros_like(flat_p)
        bn_state = {k: v.clone() for k, v in ref_model.state_dict().items() if "running" in k or "num_batches" in k}
        for step in range(2):
            grads = []
            for r in range(world):
                ref_model.load_state_dict(bn_state, strict=False)        # every rank's BatchNorm buffers start from the same state
                ref_model.dropout_masks = masks_of(r)
                for p in params:
                    p.grad = None
                batches[r].x.grad = None
                T.step_loss(ref_model, batches[r], bench.LAMBDA, True).backward()
                ref_model._pe_cache = None
                ref_model._w_cache = None
                grads.append(torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params]).double())
                if r == 0:
                    bn_next = {k: v.clone() for k, v in ref_model.state_dict().items() if "running" in k or "num_batches" in k}
            bn_state = bn_next                                           # rank 0's buffers (what rank 0 of the DP run holds)
            gavg = sum(grads) / world
            if step == 0:
                g_first = gavg.clone()
            t = step + 1
            m1 = 0.9 * m1 + 0.1 * gavg
            v1 = 0.999 * v1 + 0.001 * gavg * gavg
            flat_p = flat_p - 1e-3 / (1 - 0.9 ** t) * m1 / (v1.sqrt() / (1 - 0.999 ** t) ** 0.5 + 1e-8)
            off = 0
            with torch.no_grad():
                for p in params:
                    p.copy_(flat_p[off:off + p.numel()].view_as(p).float())
                    off += p.numel()
        # compare parameter by parameter (FlatAdam pads every parameter to 4 floats)
        # coordinates whose gradient vanishes (e.g. the attention's key bias) follow the sign of rounding noise under Adam: skipped
        worst, off = 0.0, 0
        for p, o in zip(params, opt.offsets):
            a = dp_params[o:o + p.numel()].double().cpu()
            b = flat_p[off:off + p.numel()].cpu()
            gf = g_first[off:off + p.numel()].abs().cpu()
            live = gf > 1e-4 * max(float(gf.max()), 1e-30)
            if bool(live.any()):
                scale = max(float(b.abs().max()), 1e-6)
                worst = max(worst, float((a - b).abs()[live].max()) / scale)
            off += p.numel()
        res["dp_vs_single_process_worst_rel"] = worst
    # replicas identical
    chk = torch.stack([dp_params.double().sum(), dp_params.double().abs().sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    res["replicas_identical"] = bool(torch.equal(lo, hi))
    torch.save(res, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)          # see bench.finish(): tearing NCCL down after symmetric-memory use can hang at exit


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_matches_documented_rule(tmp_path):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0, "rank process failed (exit code %s)" % p.exitcode
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    # (1) the fused kernel sums in rank order, NCCL in its own order: equal to rounding
    assert r0["allreduce_adam_max_abs_diff"] < 1e-6 and r1["allreduce_adam_max_abs_diff"] < 1e-6, (r0, r1)
    assert r0["allreduce_adam_moment_diff"] < 1e-6
    assert r0["replicas_identical"] and r1["replicas_identical"]
    # (2) two data-parallel steps == per-shard gradients averaged, per-rank BatchNorm / consistency loss
    assert r0["dp_vs_single_process_worst_rel"] < 2e-4, r0
