"""GPU parity of the small fused pieces (csrc/glue.cu) against plain torch fp32/fp64 references of the same ops:
BatchNorm1d+ReLU+dropout-mask heads (kernel/go_model.py:117-146), loss_probability (kernel/sgcn_img_snp.py:153-181) and the
Laplacian quadratic form of consist_loss (kernel/sgcn_img_snp.py:183-196).  Tolerance 1e-4 relative (north star)."""
import types

import pytest
import torch
import torch.nn.functional as F

from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("shape,groups,with_mask", [((64, 19, 32), 2, False), ((64, 19), 2, True), ((48, 54), 1, True),
                                                    ((512, 32), 2, True), ((6, 5, 3), 3, False)])
def test_bn_act_vs_torch(shape, groups, with_mask):
    from igcn_b200 import ops
    g = torch.Generator().manual_seed(0)
    z = (torch.randn(shape, generator=g) * 2 + 0.5).to(DEV)
    C = shape[1]
    bn = torch.nn.BatchNorm1d(C).to(DEV).train()
    ref = torch.nn.BatchNorm1d(C).to(DEV).double().train()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.rand(C, generator=g) - 0.5)
        ref.weight.copy_(bn.weight.double())
        ref.bias.copy_(bn.bias.double())
    mask = ((torch.rand(shape, generator=g) > 0.5).float() * 2).to(DEV) if with_mask else None
    z1 = z.clone().requires_grad_(True)
    y = ops.bn_act(z1, bn, mask, groups)
    w = torch.randn(shape, generator=g).to(DEV)
    (y * w).sum().backward()
    z2 = z.double().clone().requires_grad_(True)
    h = shape[0] // groups
    yr = torch.cat([F.relu(ref(z2[i * h:(i + 1) * h])) for i in range(groups)], 0)
    if mask is not None:
        yr = yr * mask.double()
    (yr * w.double()).sum().backward()
    H.assert_close(y, yr, what="bn_act y")
    # dz goes through a cancellation (g - mean(g) - xhat*mean(g*xhat)): rule B of tests/helpers.py with torch's own fp32
    # BatchNorm as the fp32 reference and the fp64 module as the truth
    bn32 = torch.nn.BatchNorm1d(C).to(DEV).train()
    bn32.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in ref.state_dict().items()})
    with torch.no_grad():
        bn32.running_mean.zero_()
        bn32.running_var.fill_(1.0)
    z3 = z.clone().requires_grad_(True)
    y3 = torch.cat([F.relu(bn32(z3[i * h:(i + 1) * h])) for i in range(groups)], 0)
    if mask is not None:
        y3 = y3 * mask
    (y3 * w).sum().backward()
    H.assert_parity(z1.grad, z3.grad, z2.grad, what="bn_act dz %s" % (shape,))
    H.assert_close(bn.weight.grad, ref.weight.grad, what="bn_act dgamma")
    H.assert_close(bn.bias.grad, ref.bias.grad, what="bn_act dbeta")
    H.assert_close(bn.running_mean, ref.running_mean, what="running_mean")
    H.assert_close(bn.running_var, ref.running_var, what="running_var")
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked) == groups


@pytest.mark.parametrize("N,C,K,L,with_mask", [(512, 19, 5, 32, False), (512, 19, 5, 1, True), (512, 54, 2, 1, True), (64, 7, 8, 32, True),
                                               (30, 5, 3, 4, False)])
def test_lin_bn_act_vs_torch(N, C, K, L, with_mask):
    """The read-out Linear fused with its BatchNorm head (igcn_lin_bn_act_*; kernel/go_model.py:117-131) against fp64 torch:
    nn.Linear(bias=False) -> BatchNorm1d (training, one call per stacked pass) -> relu -> mask.  The last shape (L = 4) is outside the fused
    kernel's range and goes through the unfused ops (same checks)."""
    from igcn_b200 import _lib, ops
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(N, C, K, generator=g) + 0.3).to(DEV)
    W = (torch.randn(L, K, generator=g) * 0.7).to(DEV)
    bn = torch.nn.BatchNorm1d(C).to(DEV).train()
    ref = torch.nn.BatchNorm1d(C).to(DEV).double().train()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.rand(C, generator=g) - 0.5)
        ref.weight.copy_(bn.weight.double())
        ref.bias.copy_(bn.bias.double())
    mshape = (N, C) if L == 1 else (N, C, L)
    mask = ((torch.rand(mshape, generator=g) > 0.4).float() / 0.6).to(DEV) if with_mask else None
    assert bool(_lib.lib().igcn_lin_bn_act_supported(N, C, L, K, 2)) == (L in (1, 32))
    x1, W1 = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
    y = ops.lin_bn_act(x1, W1, bn, mask, groups=2)
    w = torch.randn(N, C, L, generator=g).to(DEV)
    (y * w).sum().backward()
    x2, W2 = x.double().clone().requires_grad_(True), W.double().clone().requires_grad_(True)
    z = x2 @ W2.t()
    h = N // 2
    zz = z.squeeze(-1) if L == 1 else z
    yr = torch.cat([F.relu(ref(zz[i * h:(i + 1) * h])) for i in range(2)], 0)
    if mask is not None:
        yr = yr * mask.double()
    yr = yr.reshape(N, C, L)
    (yr * w.double()).sum().backward()
    H.assert_close(y, yr, what="lin_bn_act y")
    # fp32 reference for rule B: the same torch ops in fp32
    x3, W3 = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
    bn32 = torch.nn.BatchNorm1d(C).to(DEV).train()
    with torch.no_grad():
        bn32.weight.copy_(ref.weight.float())
        bn32.bias.copy_(ref.bias.float())
    z3 = x3 @ W3.t()
    z3 = z3.squeeze(-1) if L == 1 else z3
    y3 = torch.cat([F.relu(bn32(z3[i * h:(i + 1) * h])) for i in range(2)], 0)
    if mask is not None:
        y3 = y3 * mask
    (y3.reshape(N, C, L) * w).sum().backward()
    H.assert_parity(x1.grad, x3.grad, x2.grad, what="lin_bn_act dx %s" % ((N, C, K, L),))
    H.assert_parity(W1.grad, W3.grad, W2.grad, what="lin_bn_act dW %s" % ((N, C, K, L),))
    H.assert_parity(bn.weight.grad, bn32.weight.grad, ref.weight.grad, what="lin_bn_act dgamma")
    H.assert_parity(bn.bias.grad, bn32.bias.grad, ref.bias.grad, what="lin_bn_act dbeta")
    H.assert_close(bn.running_mean, ref.running_mean, what="lin_bn_act running_mean")
    H.assert_close(bn.running_var, ref.running_var, what="lin_bn_act running_var")
    assert int(bn.num_batches_tracked) == 2


@pytest.mark.parametrize("R,S,E", [(90, 54, 7000), (264, 10000, 300001), (5, 3, 0)])
def test_mask_loss_vs_torch(R, S, E):
    from igcn_b200 import ops
    g = torch.Generator().manual_seed(1)
    hp = types.SimpleNamespace(lamda_x_l1=0.1, lamda_e_l1=0.2, lamda_x_ent=0.3, lamda_e_ent=0.05)
    prob = (torch.randn(R, 3, generator=g)).to(DEV)
    snps = (torch.randn(1, S, generator=g)).to(DEV)
    pe = torch.rand(E, generator=g).to(DEV)
    a = [t.clone().requires_grad_(True) for t in (prob, pe, snps)]
    loss = ops.mask_loss(a[0], a[1], a[2], hp)
    (loss * 1.7).backward()

    def l1_en(p, eps=1e-6):
        n = max(p.numel(), 1)
        return p.abs().sum() / n, -(p * torch.log(p + eps) + (1 - p) * torch.log(1 - p + eps)).sum() / n
    b = [t.double().clone().requires_grad_(True) for t in (prob, pe, snps)]
    f1, fe = l1_en(torch.sigmoid(b[0]))
    e1, ee = l1_en(b[1])
    s1, se = l1_en(torch.sigmoid(b[2]))
    ref = hp.lamda_x_l1 * (f1 + s1) + hp.lamda_e_l1 * e1 + hp.lamda_x_ent * (fe + se) + hp.lamda_e_ent * ee
    (ref * 1.7).backward()
    H.assert_close(loss, ref, what="mask_loss")
    for x, y, n in zip(a, b, ("prob", "p_e", "snps_prob")):
        if y.numel():
            H.assert_close(x.grad, y.grad, what="mask_loss d" + n)


@pytest.mark.parametrize("B,D", [(64, 2880), (37, 101), (256, 8448)])
def test_laplacian_quadratic_vs_reference_form(B, D):
    """Value and gradient of tr(s^T (Dg - W) s)/B^2 as the reference writes it (fp64) vs the one-product form."""
    from igcn_b200 import ops
    g = torch.Generator().manual_seed(2)
    s = torch.rand(B, D, generator=g).to(DEV)
    t = torch.rand(B, 30, generator=g).to(DEV)
    W, d = ops.rbf_similarity(t, 0.01)                    # own kernels (csrc/laplacian.cu) vs torch in fp64
    Wd = torch.exp(-0.01 * torch.cdist(t.double(), t.double()) ** 2)
    H.assert_close(W, Wd, what="rbf similarity")
    H.assert_close(d, Wd.sum(1), what="rbf row sums")
    s1 = s.clone().requires_grad_(True)
    v = ops.laplacian_quadratic(s1, W, d, 1.0 / (B * B))
    (v * 3.0).backward()
    s2 = s.double().clone().requires_grad_(True)
    L = torch.eye(B, device=DEV, dtype=torch.float64) * Wd.sum(1) - Wd
    ref = torch.trace(s2.t() @ L @ s2) / (B * B)
    (ref * 3.0).backward()
    H.assert_close(v, ref, what="quadratic form")
    H.assert_close(s1.grad, s2.grad, what="quadratic form gradient")
    # the all-ones similarity of isSoftSimilarity=False (kernel/sgcn_img_snp.py:190): no product at all
    s3 = s.clone().requires_grad_(True)
    v1 = ops.laplacian_quadratic(s3, None, None, 1.0 / (B * B))
    v1.backward()
    s4 = s.double().clone().requires_grad_(True)
    one = torch.ones(B, B, device=DEV, dtype=torch.float64)
    r1 = torch.trace(s4.t() @ (torch.diag(one.sum(1)) - one) @ s4) / (B * B)
    r1.backward()
    H.assert_close(v1, r1, what="quadratic form, ones")
    H.assert_close(s3.grad, s4.grad, what="quadratic form gradient, ones")


def test_laplacian_quadratic_two_halves():
    from igcn_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, D = 48, 320
    s = torch.rand(2 * B, D, generator=g).to(DEV)
    t = torch.rand(B, 30, generator=g).to(DEV)
    W, d = ops.rbf_similarity(t, 0.01)
    s1 = s.clone().requires_grad_(True)
    v = ops.laplacian_quadratic(s1, W, d, 1.0 / (B * B), halves=2)
    v.backward()
    s2 = s.double().clone().requires_grad_(True)
    Wd = torch.exp(-0.01 * torch.cdist(t.double(), t.double()) ** 2)
    L = torch.diag(Wd.sum(1)) - Wd
    ref = (torch.trace(s2[:B].t() @ L @ s2[:B]) + torch.trace(s2[B:].t() @ L @ s2[B:])) / (B * B)
    ref.backward()
    H.assert_close(v, ref, what="paired quadratic form")
    H.assert_close(s1.grad, s2.grad, what="paired quadratic form gradient")


@pytest.mark.parametrize("rows_shape,Kin,Lout", [((512, 19), 5, 32), ((512, 19), 5, 1), ((64, 54), 2, 1), ((3, 7), 8, 64), ((1000,), 3, 20), ((512,), 19, 32), ((512,), 32, 32), ((77,), 32, 64)])
def test_skinny_linear_vs_torch(rows_shape, Kin, Lout):
    from igcn_b200 import ops
    g = torch.Generator().manual_seed(7)
    x = torch.randn(rows_shape + (Kin,), generator=g).to(DEV)
    W = torch.randn(Lout, Kin, generator=g).to(DEV)
    go = torch.randn(rows_shape + (Lout,), generator=g).to(DEV)
    x1, W1 = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
    z = ops.skinny_linear(x1, W1)
    (z * go).sum().backward()
    x2, W2 = x.double().requires_grad_(True), W.double().requires_grad_(True)
    zr = x2 @ W2.t()
    (zr * go.double()).sum().backward()
    H.assert_close(z, zr, what="skinny z")
    H.assert_close(x1.grad, x2.grad, what="skinny dx")
    H.assert_close(W1.grad, W2.grad, what="skinny dW")


def test_snp_mask_pair_vs_torch():
    from igcn_b200 import ops
    g = torch.Generator().manual_seed(11)
    snps = (torch.randint(0, 3, (37, 54), generator=g).float() * 0.5).to(DEV)
    p = torch.randn(1, 54, generator=g).to(DEV)
    p1 = p.clone().requires_grad_(True)
    out = ops.snp_mask_pair(snps, p1)
    w = torch.randn(74, 54, generator=g).to(DEV)
    (out * w).sum().backward()
    p2 = p.double().clone().requires_grad_(True)
    ref = torch.cat([snps.double(), snps.double() * torch.sigmoid(p2)], 0)
    (ref * w.double()).sum().backward()
    H.assert_close(out, ref, what="snp_mask_pair")
    H.assert_close(p1.grad, p2.grad, what="d snps_prob")


@pytest.mark.parametrize("rows,with_masks,use_logp", [(512, True, True), (300, False, True), (130, True, False)])
def test_output_heads_vs_torch(rows, with_masks, use_logp):
    from igcn_b200 import ops
    g = torch.Generator().manual_seed(13)
    lin2, lin2r = torch.nn.Linear(64, 3).to(DEV), torch.nn.Linear(64, 3).to(DEV)
    h1, h2 = torch.randn(rows, 64, generator=g).to(DEV), torch.randn(rows, 64, generator=g).to(DEV)
    m1 = ((torch.rand(rows, 64, generator=g) > 0.5).float() * 2).to(DEV) if with_masks else None
    m2 = ((torch.rand(rows, 64, generator=g) > 0.3).float() / 0.7).to(DEV) if with_masks else None
    a1, a2 = h1.clone().requires_grad_(True), h2.clone().requires_grad_(True)
    logp, reg = ops.output_heads(a1, m1, a2, m2, lin2, lin2r)
    w1, w2 = torch.randn(rows, 3, generator=g).to(DEV), torch.randn(rows, 3, generator=g).to(DEV)
    ((logp * w1).sum() * (1.0 if use_logp else 0.0) + (reg * w2).sum()).backward() if use_logp else (reg * w2).sum().backward()
    r2, r2r = torch.nn.Linear(64, 3).to(DEV).double(), torch.nn.Linear(64, 3).to(DEV).double()
    r2.load_state_dict({k: v.double() for k, v in lin2.state_dict().items()})
    r2r.load_state_dict({k: v.double() for k, v in lin2r.state_dict().items()})
    b1, b2 = h1.double().requires_grad_(True), h2.double().requires_grad_(True)
    x1 = b1 * m1.double() if with_masks else b1
    x2 = b2 * m2.double() if with_masks else b2
    lr, rr = F.log_softmax(r2(x1), -1), r2r(x2)
    ((lr * w1.double()).sum() * (1.0 if use_logp else 0.0) + (rr * w2.double()).sum()).backward()
    H.assert_close(logp, lr, what="logp")
    H.assert_close(reg, rr, what="reg")
    H.assert_close(a2.grad, b2.grad, what="dh2")
    H.assert_close(lin2r.weight.grad, r2r.weight.grad, what="dW2r")
    H.assert_close(lin2r.bias.grad, r2r.bias.grad, what="db2r")
    if use_logp:
        H.assert_close(a1.grad, b1.grad, what="dh1")
        H.assert_close(lin2.weight.grad, r2.weight.grad, what="dW2")
        H.assert_close(lin2.bias.grad, r2.bias.grad, what="db2")


def test_step_loss_pair_vs_torch():
    from igcn_b200 import ops
    g = torch.Generator().manual_seed(17)
    B, S = 50, 54
    reg2, cs = torch.randn(2 * B, 3, generator=g).to(DEV), torch.rand(B * 3, generator=g).to(DEV)
    xh2, snps = torch.randn(2 * B, S, generator=g).to(DEV), torch.rand(B, S, generator=g).to(DEV)
    lp, q = torch.rand((), generator=g).to(DEV), torch.rand((), generator=g).to(DEV)
    a = [t.clone().requires_grad_(True) for t in (reg2, xh2, lp, q)]
    loss = ops.step_loss_pair(a[0], cs, a[1], snps, a[2], a[3], 1.0, 1.5e-6 / 2, 0.5, 0.05)
    (loss * 2.0).backward()
    b = [t.double().clone().requires_grad_(True) for t in (reg2, xh2, lp, q)]
    csd, sd = cs.double(), snps.double()
    ref = 1.0 * (F.mse_loss(b[0][:B].reshape(-1), csd) + F.mse_loss(b[0][B:].reshape(-1), csd)) / 2 + 0.5 * b[2] \
        + 1.5e-6 * (((b[1][:B] - sd) ** 2).sum() + ((b[1][B:] - sd) ** 2).sum()) / 2 + 0.05 * b[3]
    (ref * 2.0).backward()
    H.assert_close(loss, ref, what="step loss")
    for x, y, n in zip(a, b, ("reg", "xhat", "loss_prob", "quad")):
        H.assert_close(x.grad, y.grad, what="d " + n)


def test_dp_allreduce_adam_single_rank_equals_adam_step():
    """The fused all-reduce + Adam kernel with world = 1 (flags exchanged with itself) must equal igcn_adam_step bit for bit;
    the multi-rank protocol is exercised by `bench.py --gpus N` (it reports `replicas_identical`)."""
    import ctypes
    from igcn_b200 import _lib
    g = torch.Generator().manual_seed(23)
    n = 4 * 1031
    p0 = torch.randn(n, generator=g).to(DEV)
    grad = torch.randn(n, generator=g).to(DEV)
    m0, v0 = torch.rand(n, generator=g).to(DEV) * 0.1, torch.rand(n, generator=g).to(DEV) * 0.01
    step = torch.full((1,), 3.0, device=DEV)
    lr = torch.full((1,), 1e-3, device=DEV)
    pad_bytes = 2048
    pad = torch.zeros(pad_bytes // 4, dtype=torch.int32, device=DEV)
    pa, ma, va = p0.clone(), m0.clone(), v0.clone()
    pb, mb, vb = p0.clone(), m0.clone(), v0.clone()
    gp, sp = (ctypes.c_int64 * 1)(grad.data_ptr()), (ctypes.c_int64 * 1)(pad.data_ptr())
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    for _ in range(3):                                   # replays: the flags must reset themselves
        _lib.call("igcn_dp_allreduce_adam", ctypes.addressof(gp), ctypes.addressof(sp), 0, 1, pad_bytes, _lib.ptr(pa), _lib.ptr(ma),
                  _lib.ptr(va), _lib.ptr(step), _lib.ptr(lr), 0.9, 0.999, 1e-8, n, 1000, _lib.ptr(err), _lib.stream())
        _lib.call("igcn_adam_step", _lib.ptr(pb), _lib.ptr(grad), _lib.ptr(mb), _lib.ptr(vb), _lib.ptr(step), _lib.ptr(lr), 0.9, 0.999,
                  1e-8, 1.0, n, _lib.stream())
        step += 1.0
    torch.cuda.synchronize()
    assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb)
    assert int(pad.abs().sum()) == 0 and int(err) == 0


def test_gather_flat_bit_exact():
    """FlatAdam.gather_grads (igcn_gather_flat): every p.grad lands in its slot bit for bit, parameters without a gradient read
    zero, slot padding stays zero -- more tensors than one launch's table holds, 16-byte-unaligned sources, sizes 1 .. 70 001."""
    from igcn_b200.train import FlatAdam
    g = torch.Generator().manual_seed(3)
    sizes = [1, 3, 4, 5, 31, 1024, 70001, 4099] + [int(v) for v in torch.randint(1, 3000, (120,), generator=g)]
    params = [torch.nn.Parameter(torch.randn(n, generator=g).to(DEV)) for n in sizes]
    opt = FlatAdam(params, lr=1e-3)
    pool = torch.randn(sum(sizes) + len(sizes), generator=g).to(DEV)
    off = 1                                           # +1: sources that are views at odd element offsets (not 16-byte aligned)
    for i, (p, n) in enumerate(zip(params, sizes)):
        if i % 7 == 3:
            p.grad = None
        elif i % 2:
            p.grad = pool[off:off + n]
        else:
            p.grad = torch.randn(n, generator=g).to(DEV)
        off += n
    opt.flat_grad.fill_(7.0)                          # stale contents must not survive in any slot
    pad = torch.ones(opt.n, dtype=torch.bool, device=DEV)
    for o, n in zip(opt.offsets, sizes):
        pad[o:o + n] = False
    opt.flat_grad[pad] = 0.0
    opt.gather_grads()
    torch.cuda.synchronize()
    for p, v in zip(params, opt.grad_views):
        want = torch.zeros_like(p) if p.grad is None else p.grad
        assert torch.equal(v, want)
    assert float(opt.flat_grad[pad].abs().sum()) == 0.0
