#!/usr/bin/env python
"""Runs the igcn kernels in isolation (for `ncu -k regex:igcn`) and prints CUDA-event timings + roofline fractions.

    python tools/prof_kernels.py --B 4096 --R 264 [--iters 5] [--what sgcn|attn|tc|go|gat|all]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--R", type=int, default=264)
    ap.add_argument("--L", type=int, default=2)
    ap.add_argument("--H", type=int, default=16)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--what", default="sgcn")
    ap.add_argument("--pool", default="20,15,10,8,1")
    ap.add_argument("--S", type=int, default=54)
    ap.add_argument("--compact", action="store_true")
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from igcn_b200 import _lib, ops, synthetic as syn
    from igcn_b200.data import Batch, SubjectSet
    dev = torch.device("cuda", 0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    res = {}
    if a.what in ("sgcn", "all"):
        nuniq = min(a.B, 256)                       # generate 256 distinct subjects, tile them to B (host generation cost)
        sub = syn.make_subjects(nuniq, rois=a.R, n_snps=a.S, seed=7)
        idx = np.arange(a.B) % nuniq
        b = Batch.collate(SubjectSet(sub), idx, dev)
        g = torch.Generator().manual_seed(0)
        Ws = [((torch.rand(a.H, 3 if l == 0 else a.H, generator=g) - 0.5)).to(dev).requires_grad_(True) for l in range(a.L)]
        bs = [((torch.rand(a.H, generator=g) - 0.5) * 0.1).to(dev).requires_grad_(True) for l in range(a.L)]
        prob = (torch.rand(a.R, 3, generator=g) - 0.5).to(dev).requires_grad_(True)
        pb = (torch.rand(6, 1, generator=g) - 0.5).to(dev).requires_grad_(True)
        x = b.x.clone().requires_grad_(True)
        E, N, LH = b.csr.E, a.B * a.R, a.L * a.H
        for explain in (False, True):
            go = None
            for it in range(a.iters + 2):
                flush.zero_()
                if it == 2:
                    _lib.profile_begin()
                out, pe = ops.sgcn_encoder(x, b.csr, Ws, bs, prob if explain else None, pb if explain else None, want_pe=explain)
                if go is None:
                    go = torch.randn_like(out)
                flush.zero_()
                out.backward(go)
            prof = _lib.profile_end()
            for k, (c, tot, _nb) in prof.items():
                ms = tot / c
                if "fwd" in k:
                    ab = N * 12 + E * 8 + (N + 1) * 4 + N * LH * 4 + (E * 4 if explain else 0)
                else:
                    ab = 2 * N * LH * 4 + N * 12 + E * 8 + (N + 1) * 8 + E * 4 + N * 12
                res[k] = dict(us=ms * 1e3, alg_MB=ab / 1e6, GBs=ab / ms / 1e6, frac_of_measured_peak=ab / ms / 1e6 / peak)
    if a.what in ("attn", "all"):
        torch.manual_seed(0)
        M = 19
        mha = torch.nn.MultiheadAttention(32, 2, batch_first=True).to(dev)
        q = torch.randn(a.B, a.R, 32, device=dev, requires_grad=True)
        kv = torch.randn(a.B, M, 32, device=dev, requires_grad=True)
        go = torch.randn(a.B, a.R, 32, device=dev)
        for it in range(a.iters + 2):
            flush.zero_()
            if it == 2:
                _lib.profile_begin()
            out = ops.cross_attention(q, kv, mha, relu=True)
            flush.zero_()
            out.backward(go)
        for k, (c, tot, nb) in _lib.profile_end().items():
            ms = tot / c
            res[k] = dict(us=ms * 1e3, alg_MB=nb / 1e6, GBs=nb / ms / 1e6, frac_of_measured_peak=nb / ms / 1e6 / peak)
    if a.what in ("gat", "all"):
        # BASELINE configs[1] (SGCN_GCN): GATConv(3 -> H, edge_dim=1) + GATConv(H -> H) over a batch of brain graphs
        from igcn_b200 import pyg
        nuniq = min(a.B, 256)
        sub = syn.make_subjects(nuniq, rois=a.R, n_snps=a.S, seed=7)
        b = Batch.collate(SubjectSet(sub), np.arange(a.B) % nuniq, dev)
        torch.manual_seed(0)
        convs = [pyg.GATConv(3, a.H, edge_dim=1).to(dev), pyg.GATConv(a.H, a.H, edge_dim=1).to(dev)]
        x = b.x.clone().requires_grad_(True)
        ea = b.edge_attr.clone().requires_grad_(True)
        go = None
        for it in range(a.iters + 2):
            flush.zero_()
            if it == 2:
                _lib.profile_begin()
            h = convs[1](torch.relu(convs[0](x, b.edge_index, ea)), b.edge_index, ea)
            if go is None:
                go = torch.randn_like(h)
            h.backward(go)
        for k, (c, tot, nb) in _lib.profile_end().items():
            ms = tot / c
            res[k] = dict(us=ms * 1e3, alg_MB=nb / 1e6, GBs=nb / ms / 1e6, frac_of_measured_peak=nb / ms / 1e6 / peak)
    if a.what in ("tc", "all"):
        # the Laplacian product of consist_loss at the given batch (B x B)(B x D), D = R * L * H, both passes (N = 2D)
        D = a.R * a.L * a.H
        g = torch.Generator().manual_seed(0)
        s2 = torch.rand(2 * a.B, D, generator=g).to(dev).requires_grad_(True)
        t = torch.rand(a.B, a.R, generator=g).to(dev)
        for it in range(a.iters + 2):
            flush.zero_()
            if it == 2:
                _lib.profile_begin()
            W, dd = ops.rbf_similarity(t, 0.01)
            v = ops.laplacian_quadratic(s2, W, dd, 1.0 / (a.B * a.B), halves=2)
            v.backward()
        for k, (c, tot, nb) in _lib.profile_end().items():
            ms = tot / c
            res[k] = dict(us=ms * 1e3)
            if k.startswith("laplacian_product_tc"):
                fl = 2.0 * a.B * a.B * 2 * D
                res[k].update(fp32_equiv_TFLOPs=fl / ms / 1e9, tf32_TFLOPs=3 * fl / ms / 1e9)
    if a.what in ("go", "all"):
        from igcn_b200.go_net import Gene_ontology_network
        pool = [int(v) for v in a.pool.split(",")]
        adj, go_snps, pool_dim = syn.make_go_hierarchy(pool, a.S, seed=0)
        A = torch.tensor(adj).float().t().to_sparse().coalesce()
        A_g = torch.tensor(go_snps).float().to_sparse().coalesce()
        net = Gene_ontology_network(A_g, A, 2, 2, [5, 5], pool_dim, 32, dev, dim_snps_atten=32).to(dev).train()
        data = (torch.randint(0, 3, (a.B, a.S), device=dev).float() * 0.5).requires_grad_(True)
        for it in range(a.iters + 2):
            flush.zero_()
            if it == 2:
                _lib.profile_begin()
            lat, xd, _, att = net(data)
            (lat.sum() + xd.sum() + att.sum()).backward()
        prof = _lib.profile_end()
        for k, (c, tot, _nb) in prof.items():
            res[k] = dict(us=tot / c * 1e3)
    if a.compact:
        for k, v in res.items():
            extra = ("%.3f of peak" % v["frac_of_measured_peak"]) if "frac_of_measured_peak" in v else ""
            if "tf32_TFLOPs" in v:
                extra = "%.0f TFLOP/s fp32-equivalent, %.0f TF32 TFLOP/s issued" % (v["fp32_equiv_TFLOPs"], v["tf32_TFLOPs"])
            print("%-44s %9.1f us  %s" % (k, v["us"], extra))
    else:
        print(json.dumps(dict(B=a.B, R=a.R, L=a.L, H=a.H, peak_GBs=peak, kernels=res), indent=1))


if __name__ == "__main__":
    main()
