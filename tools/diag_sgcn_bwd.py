"""Diagnostic (GPU): SGCN encoder backward, tensor-core kernels vs the register-tiled kernels, per graph."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from igcn_b200 import ops, synthetic as syn
from igcn_b200.data import Batch, SubjectSet

dev = torch.device("cuda", 0)
def run(n, R, explain, mma):
    os.environ["IGCN_SGCN_MMA"] = "1" if mma else "0"
    sub = syn.make_subjects(min(n, 64), rois=R, n_snps=8, seed=3)
    b = Batch.collate(SubjectSet(sub), np.arange(n) % min(n, 64), dev)
    g = torch.Generator().manual_seed(0)
    Ws = [(torch.rand(16, 3, generator=g) - 0.5).to(dev).requires_grad_(True), (torch.rand(16, 16, generator=g) - 0.5).to(dev).requires_grad_(True)]
    bs = [((torch.rand(16, generator=g) - 0.5) * 0.1).to(dev).requires_grad_(True) for _ in range(2)]
    prob = (torch.rand(R, 3, generator=g) - 0.5).to(dev).requires_grad_(True)
    pb = (torch.rand(6, 1, generator=g) - 0.5).to(dev).requires_grad_(True)
    x = b.x.clone().requires_grad_(True)
    out, pe = ops.sgcn_encoder(x, b.csr, Ws, bs, prob if explain else None, pb if explain else None, want_pe=explain)
    go = torch.randn(out.shape, generator=g).to(dev)
    out.backward(go)
    torch.cuda.synchronize()
    return out.detach(), x.grad.view(n, R, 3), [w.grad for w in Ws], [t.grad for t in bs]

for n, R in ((333, 90), (149, 90), (300, 90), (450, 264)):
    for explain in (False, True):
        o1, dx1, dw1, db1 = run(n, R, explain, True)
        o0, dx0, dw0, db0 = run(n, R, explain, False)
        err = (dx1 - dx0).abs().amax(dim=(1, 2)) / dx0.abs().amax()
        bad = torch.nonzero(err > 1e-4).view(-1).tolist()
        print("n=%d R=%d explain=%s: out diff %.2e; graphs with wrong dx: %d %s ; dW2 rel %.2e db2 rel %.2e dW1 rel %.2e" % (
            n, R, explain, float((o1 - o0).abs().max()), len(bad), bad[:24],
            float((dw1[1] - dw0[1]).abs().max() / dw0[1].abs().max()), float((db1[1] - db0[1]).abs().max() / db0[1].abs().max()),
            float((dw1[0] - dw0[0]).abs().max() / dw0[0].abs().max())))
        if bad:
            gb = bad[0]
            rows = torch.nonzero((dx1[gb] - dx0[gb]).abs().amax(1) > 1e-4 * dx0.abs().amax()).view(-1).tolist()
            print("   first bad graph %d: wrong rows %s" % (gb, rows[:40]))
