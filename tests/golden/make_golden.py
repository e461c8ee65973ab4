#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the reference's own UNMODIFIED files on CPU.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

What is executed from the reference, as-is: kernel/sgcn_img_snp.py (SGCN_GCN_IMGSNP),
kernel/sgcn.py (SGCN_GCN, SGCN_GAT), kernel/go_model.py (Gene_ontology_network), batch.py +
dataloader.py (collation), sgcn_hyperparameters.py, and `train()` from
kernel/train_eval_sgcn_img_snps.py:511-548.  The third-party torch_geometric / torch_scatter
calls underneath resolve to oracle/shim (restated 2.0.2 semantics).

Dropout: torch.nn.functional.dropout / dropout2d are wrapped (torch-side, the reference files
are untouched) so every mask the reference draws is RECORDED, in call order, as a multiplicative
scale tensor; the oracle and the CUDA path replay the same masks, which makes train-mode
(batch-statistics BatchNorm + dropout) parity testable.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ig-gcn_b200"))
warnings.filterwarnings("ignore")

import synthetic as syn  # noqa: E402
from oracle import ref_loader  # noqa: E402

REF = ref_loader.load()
import torch.nn.functional as F  # noqa: E402


class RecordDropout:
    """Wraps F.dropout / F.dropout2d; records each mask as a scale tensor."""

    def __init__(self):
        self.masks = []

    def __enter__(self):
        self._d, self._d2 = F.dropout, F.dropout2d
        rec = self

        def dropout(input, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return input
            m = (torch.rand_like(input) >= p).to(input.dtype) / (1.0 - p)
            rec.masks.append(m.detach().clone())
            return input * m

        def dropout2d(input, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return input
            assert input.dim() == 3          # (B, nodes, d): whole GO nodes dropped per sample
            m = (torch.rand(input.shape[0], input.shape[1], 1) >= p).to(input.dtype) / (1.0 - p)
            rec.masks.append(m.detach().clone())
            return input * m

        F.dropout, F.dropout2d = dropout, dropout2d
        torch.nn.functional.dropout, torch.nn.functional.dropout2d = dropout, dropout2d
        return self

    def __exit__(self, *a):
        F.dropout, F.dropout2d = self._d, self._d2


def data_list(sub):
    out, ep = [], sub["edge_ptr"]
    for i in range(sub["x"].shape[0]):
        e0, e1 = ep[i], ep[i + 1]
        out.append(REF.Data(
            x=torch.from_numpy(sub["x"][i]),
            edge_index=torch.from_numpy(np.vstack([sub["edge_src"][e0:e1], sub["edge_dst"][e0:e1]])),
            edge_attr=torch.from_numpy(sub["edge_attr"][e0:e1]),
            y=torch.tensor([sub["y"][i]]), clust_y=torch.tensor([sub["clust_y"][i]]),
            snps_feat=torch.from_numpy(sub["snps_feat"][i:i + 1]), sbjID=torch.tensor([sub["sbjID"][i]]),
            tsne_fdim=torch.from_numpy(sub["tsne_fdim"][i:i + 1]),
            clini_score=torch.from_numpy(sub["clini_score"][i])))
    return out


def go_graph(pool, S, seed):
    adj, go_snps, pool_dim = syn.make_go_hierarchy(pool, S, seed)
    A = torch.tensor(adj).float().t().to_sparse().coalesce()        # train_eval_sgcn_img_snps.py:69
    A_g = torch.tensor(go_snps).float().to_sparse().coalesce()      # :70
    return adj, go_snps, pool_dim, A, A_g


def sd_np(model, prefix="P/", seeded=False):
    """state_dict as arrays.  seeded=True: parameters with more than 50 000 elements are first OVERWRITTEN with values drawn
    from a seeded torch CPU generator and stored as the recipe `Pseed/<name>` = [seed, bound, *shape] instead of the values
    (tests/helpers.seeded_param regenerates them), which keeps the R=264 / H=16 fixtures small."""
    out = {}
    for i, (k, v) in enumerate(model.state_dict().items()):
        if seeded and v.is_floating_point() and v.numel() > 50000:
            bound = 1.0 / np.sqrt(v.shape[-1])
            seed = 9000 + i
            with torch.no_grad():
                v.copy_(seeded_param(seed, bound, tuple(v.shape)))
            out[prefix.rstrip("/") + "seed/" + k] = np.asarray([seed, bound] + list(v.shape), dtype=np.float64)
        else:
            out[prefix + k] = v.detach().cpu().numpy().copy()
    return out


def seeded_param(seed, bound, shape):
    g = torch.Generator().manual_seed(int(seed))
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * float(bound)


def save(name, d):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote %s (%.1f KB, %d arrays)" % (path, os.path.getsize(path) / 1024, len(d)))


# --------------------------------------------------------------------------------------------
def case_collate():
    """a1: Batch.from_data_list through the reference DataLoader, incl. ragged edge counts."""
    sub = syn.make_subjects(7, rois=30, n_snps=12, seed=11)
    # make it ragged: drop a few edges of graphs 2 and 5 (an exactly-zero PPR entry does this in real data)
    keep = np.ones(sub["edge_src"].size, bool)
    ep = sub["edge_ptr"]
    keep[ep[2] + 3] = keep[ep[2] + 17] = keep[ep[5]] = False
    cnt = np.array([keep[ep[i]:ep[i + 1]].sum() for i in range(7)])
    for k in ("edge_src", "edge_dst", "edge_attr"):
        sub[k] = sub[k][keep]
    sub["edge_ptr"] = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    idx = [4, 0, 5, 2, 6]
    dl = data_list(sub)
    loader = REF.dataloader.DataLoader([dl[i] for i in idx], batch_size=len(idx), shuffle=False)
    (b,) = list(loader)
    out = {"sub/" + k: np.asarray(v) for k, v in sub.items()}
    out["idx"] = np.asarray(idx)
    for k in b.keys:
        out["out/" + k] = b[k].numpy()
    out["out/num_graphs"] = np.asarray(b.num_graphs)
    save("collate_r30", out)


def build_model(L, H, R, pool, S, num_regr=3, num_classes=3, seed=0):
    adj, go_snps, pool_dim, A, A_g = go_graph(pool, S, seed)
    torch.manual_seed(seed)
    cls = REF.sgcn_img_snp.SGCN_GCN_IMGSNP
    if S != 54:
        # the reference hard-codes 54 SNPs for snps_prob (sgcn_img_snp.py:96); build then resize, as the
        # parametric replacement does (n_snps argument)
        m = cls(L, H, A_g, A, pool_dim, 32, "cpu", rois=R, H_0=3, num_classes=num_classes, isCrossAtten=True,
                isSoftSimilarity=True, rbf_gamma=0.01, isuseProb4Regr=True, num_regr=num_regr,
                isImageOnly=False, isSNPsOnly=False)
        m.snps_prob = torch.nn.Parameter(torch.empty(1, S).uniform_(-1 / np.sqrt(S), 1 / np.sqrt(S)))
    else:
        m = cls(L, H, A_g, A, pool_dim, 32, "cpu", rois=R, H_0=3, num_classes=num_classes, isCrossAtten=True,
                isSoftSimilarity=True, rbf_gamma=0.01, isuseProb4Regr=True, num_regr=num_regr,
                isImageOnly=False, isSNPsOnly=False)
    # make zero-initialised parameters non-trivial so their gradients/paths are exercised
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias") and p.abs().sum() == 0:
                p.uniform_(-0.1, 0.1)
    return m, adj, go_snps, pool_dim


def case_model(name, L, H, R, B, pool, S, seed, skip_big_grads=False, steps=0, compact=False):
    """a2-a13: full SGCN_GCN_IMGSNP fwd (plain + explain), every loss, grads of one train() step.
    compact=True (the R=264 case): big parameters are seeded recipes, out_lin is stored as its latent tail only (the rest is
    out_z), and gradients of big parameters are stored as their row / column sums."""
    m, adj, go_snps, pool_dim = build_model(L, H, R, pool, S, seed=seed)
    sub = syn.make_subjects(B, rois=R, n_snps=S, seed=seed + 100)
    dl = data_list(sub)
    loader = REF.dataloader.DataLoader(dl, batch_size=B, shuffle=False)
    (b,) = list(loader)
    out = sd_np(m, seeded=compact)
    out.update({"sub/" + k: np.asarray(v) for k, v in sub.items()})
    out.update(adj=adj, go_snps=go_snps, pool=np.asarray(pool_dim[0]),
               cfg=np.asarray([L, H, R, B, S]))
    # classification terms switched on (default main.py:73 has [0]=0); orthogonality weight at its default 0
    # (main.py:78): the reference evaluates that term as an fp32 D x D product that is ~2e-3 off its own fp64
    # value, so it is pinned separately (see "orthogonal") instead of contaminating the step loss
    lam = [0.7, 1.0, 0.5, 1.5e-6, 0.1, 0.0]
    out["lambda_loss"] = np.asarray(lam)
    names = ["logp", "x_hat", "out_z", "out_lin", "linear_outf", "our_reg"]

    # eval-mode forward (running-stat BN, no dropout)
    m.eval()
    with torch.no_grad():
        for tag, ex in (("plain", False), ("explain", True)):
            o = m(b, 0.1, "cpu", isExplain=ex)
            for n, t in zip(names, o):
                out["eval/%s/%s" % (tag, n)] = t.numpy()[:, -32:] if (compact and n == "out_lin") else t.numpy()
    b.x.requires_grad_(False)

    # train-mode forwards with recorded dropout
    m.train()
    bn_before = {k: v.clone() for k, v in m.state_dict().items() if "running" in k}
    torch.manual_seed(seed + 5)
    with RecordDropout() as rec:
        o = m(b, 0.1, "cpu")
        masks_plain = list(rec.masks)
        rec.masks.clear()
        q = m(b, 0.1, "cpu", isExplain=True)
        masks_explain = list(rec.masks)
    for n, t in zip(names, o):
        out["train/plain/%s" % n] = t.detach().numpy()[:, -32:] if (compact and n == "out_lin") else t.detach().numpy()
    for n, t in zip(names, q):
        out["train/explain/%s" % n] = t.detach().numpy()[:, -32:] if (compact and n == "out_lin") else t.detach().numpy()
    from oracle.igcn_oracle import MODEL_MASK_NAMES
    assert len(masks_plain) == len(MODEL_MASK_NAMES) == len(masks_explain)
    for n, mp, me in zip(MODEL_MASK_NAMES, masks_plain, masks_explain):
        out["mask/plain/" + n], out["mask/explain/" + n] = mp.numpy(), me.numpy()
    out["loss_probability"] = m.loss_probability(b.x, b.edge_index, b.edge_attr, REF.hp).detach().numpy()
    out["consist_loss"] = m.consist_loss(o[2], b.tsne_fdim).detach().numpy()
    out["consist_loss_ones"] = m.consist_loss(o[2]).detach().numpy()
    out["orthogonal"] = m.OrthogonalConstraint(o[2]).detach().numpy()
    cp = m.cal_probability(b.x, b.edge_index, b.edge_attr, b.snps_feat)
    for n, t in zip(["x_feat_prob", "edge_weight_prob", "x_prob", "edge_prob", "snps_feat_prob", "snps_prob"], cp):
        out["calprob/" + n] = t.detach().numpy().copy()      # x_prob IS the parameter: copy before Adam touches it
    # restore BN running stats so the step below starts from the stored state_dict
    m.load_state_dict({**m.state_dict(), **bn_before})

    # one reference train() step: grads of every parameter and of data.x
    import importlib
    te = importlib.import_module("kernel.train_eval_sgcn_img_snps")
    opt = torch.optim.SGD(m.parameters(), lr=0.0)
    b.x.requires_grad_(False)
    b.x.grad = None
    torch.manual_seed(seed + 5)
    with RecordDropout() as rec:
        mean_loss = te.train(m, opt, loader, 0.1, lam, torch.nn.MSELoss(reduction="none"), True, "cpu")
    assert len(rec.masks) == 2 * len(MODEL_MASK_NAMES)
    # (iterating the torch DataLoader inside train() draws from the global RNG, so these masks differ
    #  from the ones recorded above: they are stored separately)
    for n, mp, me in zip(MODEL_MASK_NAMES, rec.masks[:9], rec.masks[9:]):
        out["stepmask/plain/" + n], out["stepmask/explain/" + n] = mp.numpy(), me.numpy()
    out["step/loss"] = np.asarray(mean_loss)
    for n, p in m.named_parameters():
        if p.grad is None:
            continue
        if (skip_big_grads or compact) and p.numel() > 50000:
            if compact:
                out["gradsum/rows/" + n] = p.grad.sum(1).numpy().copy()
                out["gradsum/cols/" + n] = p.grad.sum(0).numpy().copy()
            continue
        out["grad/" + n] = p.grad.numpy().copy()
    out.update({"bn_after/" + k: v.numpy().copy() for k, v in m.state_dict().items() if "running" in k})

    if steps:
        # loop-level: `steps` reference train() epochs with Adam, masks recorded per step
        assert not compact
        m.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in out.items() if k.startswith("P/")})
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=0)
        losses = []
        torch.manual_seed(seed + 9)
        for s in range(steps):
            with RecordDropout() as rec:
                losses.append(te.train(m, opt, loader, 0.1, lam, torch.nn.MSELoss(reduction="none"), True, "cpu"))
            for n, mp, me in zip(MODEL_MASK_NAMES, rec.masks[:9], rec.masks[9:]):
                out["adam/mask/%d/plain/%s" % (s, n)] = mp.numpy()
                out["adam/mask/%d/explain/%s" % (s, n)] = me.numpy()
        out["adam/losses"] = np.asarray(losses)
        for n in ("prob", "prob_bias", "snps_prob", "conv1.lin.weight", "go_network.t.0", "go_network.w_att_in.0.weight"):
            out["adam/final/" + n] = dict(m.named_parameters())[n].detach().numpy().copy()
    save(name, out)


def case_go(name, pool, S, B, C, seed):
    """a6-a10: Gene_ontology_network standalone (config-3 shaped, scaled down), fwd + bwd."""
    adj, go_snps, pool_dim, A, A_g = go_graph(pool, S, seed)
    torch.manual_seed(seed)
    net = REF.go_model.Gene_ontology_network(A_g, A, 2, 2, [5, 5], pool_dim, 32, "cpu", dim_snps_atten=C)
    rng = np.random.default_rng(seed)
    data = torch.from_numpy((rng.integers(0, 3, size=(B, S)) * 0.5).astype(np.float32))
    out = sd_np(net)
    out.update(adj=adj, go_snps=go_snps, pool=np.asarray(pool), data=data.numpy(), cfg=np.asarray([B, S, C]))
    for j in range(2):
        out["prep/enc%d/index" % j] = net.n_loc_in[j].numpy()
        out["prep/enc%d/store" % j] = net.store_in[j].numpy()
        out["prep/dec%d/index" % j] = net.n_loc_out[j].numpy()
        out["prep/dec%d/store" % j] = net.store_out[j].numpy()
    out["prep/ag"], out["prep/ag_t"] = net.i.numpy(), net.i_D.numpy()
    net.eval()
    with torch.no_grad():
        lat, xd, _, att = net(data, 0.1, "cpu")
    out["eval/latent"], out["eval/x_D"], out["eval/atten_out"] = lat.numpy(), xd.numpy(), att.numpy()
    net.train()
    torch.manual_seed(seed + 1)
    d = data.clone().requires_grad_(True)
    with RecordDropout() as rec:
        lat, xd, _, att = net(d, 0.1, "cpu")
    from oracle.igcn_oracle import GO_MASK_NAMES
    assert len(rec.masks) == len(GO_MASK_NAMES)
    for n, mk in zip(GO_MASK_NAMES, rec.masks):
        out["mask/" + n] = mk.numpy()
    out["train/latent"], out["train/x_D"], out["train/atten_out"] = lat.detach().numpy(), xd.detach().numpy(), att.detach().numpy()
    loss = lat.sum() + ((xd - data) ** 2).mean() + (att * torch.linspace(0.5, 1.5, att.shape[-1])).sum()
    loss.backward()
    out["train/loss"] = loss.detach().numpy()
    out["grad/data"] = d.grad.numpy()
    for n, p in net.named_parameters():
        if p.grad is not None:
            out["grad/" + n] = p.grad.numpy().copy()
    save(name, out)


def case_sgcn(name, R, B, L, H, seed, seeded=False):
    """config 1: kernel/sgcn.py SGCN_GCN (GCNConv stack) and SGCN_GAT (GATConv edge_dim=1), 3-term step."""
    sub = syn.make_subjects(B, rois=R, n_snps=4, seed=seed, num_classes=2)
    dl = data_list(sub)
    (b,) = list(REF.dataloader.DataLoader(dl, batch_size=B, shuffle=False))
    out = {"sub/" + k: np.asarray(v) for k, v in sub.items()}
    out["cfg"] = np.asarray([L, H, R, B])

    class DS:  # SGCN_GAT reads dataset.num_features / num_classes (sgcn.py:163,168)
        num_features, num_classes = 3, 2

    for tag, ctor in (("gcn", lambda: REF.sgcn.SGCN_GCN(None, L, H, rois=R)),
                      ("gat", lambda: REF.sgcn.SGCN_GAT(DS, L, H, rois=R))):
        torch.manual_seed(seed)
        m = ctor()
        # sgcn.py:285/167 hard-codes 90 ROIs in lin1; rebuild for R != 90 as the parametric replacement does
        if R != 90:
            m.lin1 = torch.nn.Linear(R * L * H, 64)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if n.endswith("bias") and p.abs().sum() == 0:
                    p.uniform_(-0.1, 0.1)
        out.update(sd_np(m, "P_%s/" % tag, seeded=seeded))
        m.eval()
        b.x.requires_grad_(False)
        b.x.grad = None
        o = m(b)
        q = m(b, True)
        lp = m.loss_probability(b.x, b.edge_index, b.edge_attr, REF.hp)
        loss = REF.hp.lamda_ce * F.nll_loss(o, b.y.view(-1)) + lp + REF.hp.lamda_mi * F.nll_loss(q, b.y.view(-1))
        loss.backward()
        out["%s/logp" % tag], out["%s/logp_explain" % tag] = o.detach().numpy(), q.detach().numpy()
        out["%s/loss_prob" % tag], out["%s/loss" % tag] = lp.detach().numpy(), loss.detach().numpy()
        out["%s/grad/x" % tag] = b.x.grad.numpy().copy()
        for n, p in m.named_parameters():
            if p.grad is not None and p.numel() <= 50000:
                out["%s/grad/%s" % (tag, n)] = p.grad.numpy().copy()
            elif p.grad is not None and seeded:
                out["%s/gradsum/rows/%s" % (tag, n)] = p.grad.sum(1).numpy().copy()
                out["%s/gradsum/cols/%s" % (tag, n)] = p.grad.sum(0).numpy().copy()
    save(name, out)


CASES = {
    "collate_r30": case_collate,
    "imgsnp_small": lambda: case_model("imgsnp_small", L=3, H=8, R=30, B=6, pool=[9, 6, 4, 3, 1], S=20, seed=1, steps=3),
    "imgsnp_adni": lambda: case_model("imgsnp_adni", L=2, H=16, R=90, B=4, pool=syn.ADNI_POOL, S=54, seed=2, skip_big_grads=True),
    "go_mid": lambda: case_go("go_mid", pool=[40, 20, 10, 5, 1], S=150, B=5, C=7, seed=3),
    "sgcn_cfg1": lambda: case_sgcn("sgcn_cfg1", R=90, B=4, L=2, H=8, seed=4),
    # BASELINE config 4's graph size (264 ROIs) through the full model, and BASELINE config 1 exactly (B=32, H=16, 90 ROIs)
    "imgsnp_r264": lambda: case_model("imgsnp_r264", L=2, H=16, R=264, B=8, pool=syn.ADNI_POOL, S=54, seed=6, compact=True),
    "sgcn_cfg1_b32": lambda: case_sgcn("sgcn_cfg1_b32", R=90, B=32, L=2, H=16, seed=7, seeded=True),
}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):
        CASES[name]()
