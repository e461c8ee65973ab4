"""Diagnostic (GPU): replicate tests/test_gpu_sgcn.py::test_sgcn_encoder_fwd_bwd[False-333-90-2-16-False] and localise the dx mismatch."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from igcn_b200 import ops
from igcn_b200.data import Batch, SubjectSet
from oracle import igcn_oracle as O
from tests.test_gpu_sgcn import _subjects, _enc_params

dev = torch.device("cuda", 0)
n, R, L, Hd = 333, 90, 2, 16
sub = _subjects(n, R, seed=3 * n + R + L, ragged=False)
idx = np.arange(n)
c = O.collate(sub, idx)
P64 = _enc_params(L, Hd, R, 3, seed=L * 100 + Hd, dtype=torch.float64)
for v in P64.values():
    v.requires_grad_(True)
x64 = torch.from_numpy(c["x"]).double().requires_grad_(True)
ei = torch.from_numpy(c["edge_index"])
w64 = torch.from_numpy(c["edge_attr"]).double()
ref = O.sgcn_encoder(P64, x64, ei, w64, L, R)
gen = torch.Generator().manual_seed(5)
g_out = torch.randn(ref.shape, generator=gen, dtype=torch.float64)
(ref * g_out).sum().backward()
res = {}
for mma in (1, 0):
    os.environ["IGCN_SGCN_MMA"] = str(mma)
    b = Batch.collate(SubjectSet(sub), idx, dev)
    Pc = {k: v.detach().float().to(dev).requires_grad_(True) for k, v in P64.items()}
    xc = b.x.clone().requires_grad_(True)
    Ws = [Pc["conv1.lin.weight"], Pc["convs.0.lin.weight"]]
    bs = [Pc["conv1.bias"], Pc["convs.0.bias"]]
    out, _ = ops.sgcn_encoder(xc, b.csr, Ws, bs)
    (out * g_out.float().to(dev)).sum().backward()
    torch.cuda.synchronize()
    res[mma] = (out.detach().cpu().double(), xc.grad.cpu().double())
for mma in (1, 0):
    o, dx = res[mma]
    d = (dx - x64.grad).abs().view(n, R, 3)
    scale = x64.grad.abs().max()
    per_graph = d.amax(dim=(1, 2)) / scale
    bad = torch.nonzero(per_graph > 1e-4).view(-1).tolist()
    print("mma=%d out err %.2e dx: %d bad graphs %s" % (mma, float((o - ref.detach()).abs().max()), len(bad), bad[:20]))
    for gb in bad[:3]:
        rows = torch.nonzero(d[gb].amax(1) / scale > 1e-4).view(-1).tolist()
        ep = sub["edge_ptr"]
        s = sub["edge_src"][ep[gb]:ep[gb + 1]]; t = sub["edge_dst"][ep[gb]:ep[gb + 1]]; w = sub["edge_attr"][ep[gb]:ep[gb + 1]]
        indeg = np.bincount(t, minlength=R); outdeg = np.bincount(s, minlength=R)
        print("   graph %d: Eg=%d wrong rows %s; indeg min/max %d/%d outdeg max %d; min |w| %.3e; self loops %d" % (
            gb, ep[gb + 1] - ep[gb], rows[:16], indeg.min(), indeg.max(), outdeg.max(), float(np.abs(w).min()), int((s == t).sum())))
        for r in rows[:4]:
            print("      row %d: got %s ref %s outdeg %d indeg %d" % (r, dx.view(n, R, 3)[gb, r].tolist(), x64.grad.view(n, R, 3)[gb, r].tolist(), outdeg[r], indeg[r]))
