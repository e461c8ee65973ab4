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
    H.assert_close(z1.grad, z2.grad, rtol=2e-4, what="bn_act dz")
    H.assert_close(bn.weight.grad, ref.weight.grad, what="bn_act dgamma")
    H.assert_close(bn.bias.grad, ref.bias.grad, what="bn_act dbeta")
    H.assert_close(bn.running_mean, ref.running_mean, what="running_mean")
    H.assert_close(bn.running_var, ref.running_var, what="running_var")
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked) == groups


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
    W = torch.exp(-0.01 * torch.cdist(t, t) ** 2)
    lap = torch.diag(W.sum(1)) - 0.5 * (W + W.t())
    s1 = s.clone().requires_grad_(True)
    v = ops.laplacian_quadratic(s1, lap, 1.0 / (B * B))
    (v * 3.0).backward()
    s2 = s.double().clone().requires_grad_(True)
    Wd = W.double()
    L = torch.eye(B, device=DEV, dtype=torch.float64) * Wd.sum(1) - Wd
    ref = torch.trace(s2.t() @ L @ s2) / (B * B)
    (ref * 3.0).backward()
    H.assert_close(v, ref, what="quadratic form")
    H.assert_close(s1.grad, s2.grad, what="quadratic form gradient")


def test_laplacian_quadratic_two_halves():
    from igcn_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, D = 48, 320
    s = torch.rand(2 * B, D, generator=g).to(DEV)
    t = torch.rand(B, 30, generator=g).to(DEV)
    W = torch.exp(-0.01 * torch.cdist(t, t) ** 2)
    lap = torch.diag(W.sum(1)) - 0.5 * (W + W.t())
    s1 = s.clone().requires_grad_(True)
    v = ops.laplacian_quadratic(s1, lap, 1.0 / (B * B), halves=2)
    v.backward()
    s2 = s.double().clone().requires_grad_(True)
    L = lap.double()
    ref = (torch.trace(s2[:B].t() @ L @ s2[:B]) + torch.trace(s2[B:].t() @ L @ s2[B:])) / (B * B)
    ref.backward()
    H.assert_close(v, ref, what="paired quadratic form")
    H.assert_close(s1.grad, s2.grad, what="paired quadratic form gradient")


@pytest.mark.parametrize("rows_shape,Kin,Lout", [((512, 19), 5, 32), ((512, 19), 5, 1), ((64, 54), 2, 1), ((3, 7), 8, 64), ((1000,), 3, 20)])
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
