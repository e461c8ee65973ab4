"""GPU parity of the tcgen05 (3xTF32) products (csrc/tc_gemm.cu) against fp64 torch: the raw product, the fusion heads
(cat_linear forward + both backward products + bias gradient) and the Laplacian quadratic form.  The tolerance is the
north-star 1e-4 relative; the observed error of the split product is ~1e-6."""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 64, 64), (5, 3, 37), (512, 64, 2912), (256, 2880, 256), (64, 3183, 512),
                                   (300, 200, 1000), (130, 129, 33), (4096, 4900, 96)])
def test_tc_matmul_vs_fp64(M, N, K):
    from igcn_b200 import ops
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    b = torch.randn(N, K, generator=g).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    out = ops.tc_matmul_nt(a, b)
    ref = a.double() @ b.double().t()
    e = H.rel_err(out, ref)
    assert e < 5e-6, "3xTF32 product rel err %.2e" % e
    out2 = ops.tc_matmul_nt(a, b, bias, relu=True)
    H.assert_close(out2, torch.relu(ref + bias.double()), rtol=5e-6, what="bias+relu epilogue")
    # run-to-run bit identical (fixed split-K summation order)
    assert torch.equal(out, ops.tc_matmul_nt(a, b))


@pytest.mark.parametrize("M,widths,N", [(64, (2880, 32, 270), 64), (7, (33, 0, 5), 64), (256, (264, 32, 0), 64)])
def test_cat_linear_tc_vs_fp64(M, widths, N):
    from igcn_b200 import ops
    if not ops.USE_TC:
        pytest.skip("tensor-core path disabled by IGCN_NO_TC")
    g = torch.Generator().manual_seed(3)
    K = sum(widths)
    xs = [torch.randn(M, w, generator=g).to(DEV).requires_grad_(True) if w else None for w in widths]
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(N, generator=g).to(DEV).requires_grad_(True)
    y = ops.cat_linear([t for t in xs], W, b, relu=True)
    go = torch.randn(M, N, generator=g).to(DEV)
    (y * go).sum().backward()
    xd = [t.detach().double().requires_grad_(True) if t is not None else None for t in xs]
    Wd, bd = W.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    yr = torch.relu(torch.cat([t for t in xd if t is not None], 1) @ Wd.t() + bd)
    (yr * go.double()).sum().backward()
    H.assert_close(y, yr, rtol=1e-5, what="cat_linear y")
    H.assert_close(W.grad, Wd.grad, rtol=1e-5, what="dW")
    H.assert_close(b.grad, bd.grad, rtol=1e-5, what="db")
    for t, r in zip(xs, xd):
        if t is not None:
            H.assert_close(t.grad, r.grad, rtol=1e-5, what="dx")
