"""Diagnostic (GPU): per-parameter gradient error of the GO network kernels vs the fp64 oracle, several seeds, B=256 ADNI shape.
Tells accumulation noise (varies with the seed) from a systematic difference (same size every time)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from igcn_b200 import synthetic as syn
from igcn_b200.go_net import Gene_ontology_network
from oracle import igcn_oracle as O
from tests import helpers as H

dev = torch.device("cuda", 0)
pool = syn.ADNI_POOL
adj, go_snps, pool_dim = syn.make_go_hierarchy(pool, 54, seed=0)
A = torch.tensor(adj).float().t().to_sparse().coalesce()
A_g = torch.tensor(go_snps).float().to_sparse().coalesce()
prep = O.go_index_prep(adj.T, go_snps, pool)
B = 512
for seed in range(4):
    torch.manual_seed(seed)
    net = Gene_ontology_network(A_g, A, 2, 2, [5, 5], pool_dim, 32, dev, dim_snps_atten=32).to(dev).train()
    rng = np.random.default_rng(seed)
    data = torch.from_numpy((rng.integers(0, 3, size=(B, 54)) * 0.5).astype(np.float32))
    shapes = dict(go_enc0=(B, 54, 1), go_enc1=(B, 34, 1), go_B=(B, 19), go_dec0=(B, 34, 1), go_dec1=(B, 54, 1), go_BD=(B, 54), go_latent=(B, 32))
    gen = torch.Generator().manual_seed(seed)
    masks = {k: (torch.rand(s, generator=gen) > 0.4).float() / 0.6 for k, s in shapes.items()}
    w = torch.linspace(0.5, 1.5, 32)
    res = {}
    for dt in (torch.float32, torch.float64):
        P = {"go_network." + k: v.detach().cpu().to(dt).requires_grad_(v.is_floating_point()) for k, v in net.state_dict().items()}
        d = data.to(dt).requires_grad_(True)
        lat, xd, att = O.go_forward(P, prep, d, True, {k: v.to(dt) for k, v in masks.items()})
        (lat.sum() + ((xd - data.to(dt)) ** 2).mean() + (att * w.to(dt)).sum()).backward()
        res[dt] = {k[len("go_network."):]: v.grad for k, v in P.items() if v.grad is not None}
    net.dropout_masks = masks
    dc = data.to(dev).requires_grad_(True)
    lat, xd, _, att = net(dc, 0.1, dev)
    (lat.sum() + ((xd - data.to(dev)) ** 2).mean() + (att * w.to(dev)).sum()).backward()
    rows = []
    for k, p in net.named_parameters():
        if p.grad is not None:
            rows.append((H.rel_err(p.grad, res[torch.float64][k]), H.rel_err(res[torch.float32][k], res[torch.float64][k]), k))
    rows.sort(reverse=True)
    print("seed", seed, " worst (ours-vs-64, oracle32-vs-64):", [("%.1e" % a, "%.1e" % b, k) for a, b, k in rows[:5]])
