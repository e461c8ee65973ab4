"""Pins oracle/igcn_oracle.py against the golden vectors produced by the reference's own,
unmodified files (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import igcn_oracle as O
from tests import helpers as H


def _batch(sub, idx=None, dtype=torch.float32):
    c = O.collate(sub, np.arange(sub["x"].shape[0]) if idx is None else idx)
    b = {k: torch.from_numpy(v) for k, v in c.items()}
    for k in ("x", "edge_attr", "snps_feat", "clini_score", "tsne_fdim"):
        b[k] = b[k].to(dtype)
    return b


def test_collate_bit_exact():
    g = H.load("collate_r30")
    sub = H.subjects(g)
    c = O.collate(sub, g["idx"])
    for k in ("x", "edge_index", "edge_attr", "batch", "snps_feat", "y", "clini_score", "tsne_fdim", "clust_y", "sbjID"):
        ref = g["out/" + k]
        assert c[k].dtype == ref.dtype, k
        assert np.array_equal(c[k], ref), k
    assert int(g["out/num_graphs"]) == len(g["idx"])


def test_go_index_prep_bit_exact():
    g = H.load("go_mid")
    prep = O.go_index_prep(g["adj"].T, g["go_snps"], list(g["pool"]))
    for j in range(2):
        assert np.array_equal(prep["enc"][j]["index"], g["prep/enc%d/index" % j])
        assert np.array_equal(prep["enc"][j]["store"], g["prep/enc%d/store" % j])
        assert np.array_equal(prep["dec"][j]["index"], g["prep/dec%d/index" % j])
        assert np.array_equal(prep["dec"][j]["store"], g["prep/dec%d/store" % j])
    assert np.array_equal(prep["ag"], g["prep/ag"])
    assert np.array_equal(prep["ag_t"], g["prep/ag_t"])


@pytest.mark.parametrize("loop", [False, True])
def test_go_network(loop):
    g = H.load("go_mid")
    prep = O.go_index_prep(g["adj"].T, g["go_snps"], list(g["pool"]))
    P = H.params(g, grad=True)
    P = {"go_network." + k: v for k, v in P.items()}
    data = torch.from_numpy(g["data"])
    with torch.no_grad():
        lat, xd, att = O.go_forward(P, prep, data, training=False, per_subject_loop=loop)
    H.assert_close(lat, g["eval/latent"], what="eval latent")
    H.assert_close(xd, g["eval/x_D"], what="eval x_D")
    H.assert_close(att, g["eval/atten_out"], what="eval atten")
    masks = {k: torch.from_numpy(v) for k, v in H.sub_dict(g, "mask/").items()}
    d = data.clone().requires_grad_(True)
    lat, xd, att = O.go_forward(P, prep, d, training=True, masks=masks, per_subject_loop=loop)
    H.assert_close(lat, g["train/latent"], what="train latent")
    H.assert_close(xd, g["train/x_D"], what="train x_D")
    H.assert_close(att, g["train/atten_out"], what="train atten")
    loss = lat.sum() + ((xd - data) ** 2).mean() + (att * torch.linspace(0.5, 1.5, att.shape[-1])).sum()
    loss.backward()
    H.assert_close(loss, g["train/loss"], what="loss")
    H.assert_close(d.grad, g["grad/data"], what="grad data")
    for k, v in H.sub_dict(g, "grad/").items():
        if k == "data":
            continue
        H.assert_close(P["go_network." + k].grad, v, what="grad " + k)


def _lin_tail(case, t, n):
    """the compact R=264 fixture stores out_lin as its latent tail (the leading part is out_z)"""
    return t[:, -32:] if (case == "imgsnp_r264" and n == "out_lin") else t


@pytest.mark.parametrize("case", ["imgsnp_small", "imgsnp_adni", "imgsnp_r264"])
def test_full_model(case):
    g = H.load(case)
    L, Hd, R, B, S = [int(v) for v in g["cfg"]]
    prep = O.go_index_prep(g["adj"].T, g["go_snps"], list(g["pool"]))
    P = H.params(g, grad=True)
    b = _batch(H.subjects(g))
    names = ["logp", "x_hat", "out_z", "out_lin", "linear_outf", "our_reg"]
    with torch.no_grad():
        for tag, ex in (("plain", False), ("explain", True)):
            o = O.model_forward(P, prep, b, L, R, ex, training=False)
            for n, t in zip(names, o):
                H.assert_close(_lin_tail(case, t, n), g["eval/%s/%s" % (tag, n)], what="eval %s %s" % (tag, n))
        for tag, ex in (("plain", False), ("explain", True)):
            masks = {k: torch.from_numpy(v) for k, v in H.sub_dict(g, "mask/%s/" % tag).items()}
            o = O.model_forward(P, prep, b, L, R, ex, training=True, masks=masks)
            for n, t in zip(names, o):
                H.assert_close(_lin_tail(case, t, n), g["train/%s/%s" % (tag, n)], what="train %s %s" % (tag, n))
            if tag == "plain":
                H.assert_close(O.consist_loss(o[2], b["tsne_fdim"], 0.01), g["consist_loss"], what="consist")
                H.assert_close(O.consist_loss(o[2]), g["consist_loss_ones"], what="consist ones")
                H.assert_close(O.orthogonal_constraint(o[2]), g["orthogonal"], what="orth")
        H.assert_close(O.loss_probability(P, b["x"], b["edge_index"], b["edge_attr"], R), g["loss_probability"], what="loss_prob")
        cp = O.cal_probability(P, b["x"], b["edge_index"], b["edge_attr"], R, b["snps_feat"])
        H.assert_close(cp["x"], g["calprob/x_feat_prob"])
        H.assert_close(cp["w"], g["calprob/edge_weight_prob"])
        H.assert_close(cp["p_e"], g["calprob/edge_prob"])
        H.assert_close(cp["snps"], g["calprob/snps_feat_prob"])
    # one train() step: loss + every gradient
    mp = {k: torch.from_numpy(v) for k, v in H.sub_dict(g, "stepmask/plain/").items()}
    me = {k: torch.from_numpy(v) for k, v in H.sub_dict(g, "stepmask/explain/").items()}
    b["x"].requires_grad_(True)
    loss, _, _ = O.train_step_loss(P, prep, b, L, R, list(g["lambda_loss"]), 0.01, True, mp, me)
    loss.backward()
    H.assert_close(loss, g["step/loss"], what="step loss")
    truth = _fp64_step_grads(g, prep, L, R, mp, me)
    for k, v in H.sub_dict(g, "grad/").items():
        assert P[k].grad is not None, k
        H.assert_parity(P[k].grad, v, truth[k], what="grad " + k)
    for k, v in H.sub_dict(g, "gradsum/rows/").items():
        H.assert_parity(P[k].grad.sum(1), v, truth[k].sum(1), what="grad row sums " + k)
    for k, v in H.sub_dict(g, "gradsum/cols/").items():
        H.assert_parity(P[k].grad.sum(0), v, truth[k].sum(0), what="grad column sums " + k)


def _fp64_step_grads(g, prep, L, R, mp, me):
    """fp64 truth of the train() step's gradients (the oracle in double precision on the same inputs and masks)."""
    P64 = H.params(g, dtype=torch.float64, grad=True)
    b64 = _batch(H.subjects(g), dtype=torch.float64)
    b64["x"].requires_grad_(True)
    d = lambda m: {k: v.double() for k, v in m.items()}
    loss, _, _ = O.train_step_loss(P64, prep, b64, L, R, list(g["lambda_loss"]), 0.01, True, d(mp), d(me))
    loss.backward()
    return {k: v.grad for k, v in P64.items() if v.grad is not None}


@pytest.mark.parametrize("case", ["sgcn_cfg1", "sgcn_cfg1_b32"])
def test_config1_sgcn_gcn(case):
    g = H.load(case)
    L, Hd, R, B = [int(v) for v in g["cfg"]]
    b = _batch(H.subjects(g))
    P = H.params(g, "P_gcn/", grad=True)
    b["x"].requires_grad_(True)
    o = O.sgcn_gcn_forward(P, b, L, R, False, False)
    q = O.sgcn_gcn_forward(P, b, L, R, True, False)
    lp = O.loss_probability_sgcn(P, b["x"], b["edge_index"], b["edge_attr"], R)
    loss = torch.nn.functional.nll_loss(o, b["y"]) + lp + torch.nn.functional.nll_loss(q, b["y"])
    loss.backward()
    H.assert_close(o, g["gcn/logp"])
    H.assert_close(q, g["gcn/logp_explain"])
    H.assert_close(lp, g["gcn/loss_prob"])
    H.assert_close(loss, g["gcn/loss"])
    H.assert_close(b["x"].grad, g["gcn/grad/x"], what="grad x")
    for k, v in H.sub_dict(g, "gcn/grad/").items():
        if k != "x":
            H.assert_close(P[k].grad, v, what="grad " + k)
    for k, v in H.sub_dict(g, "gcn/gradsum/rows/").items():
        H.assert_close(P[k].grad.sum(1), v, what="grad row sums " + k)


@pytest.mark.parametrize("case", ["sgcn_cfg1", "sgcn_cfg1_b32"])
def test_config1_gat_conv(case):
    """GATConv(edge_dim=1) layer stack of SGCN_GAT (kernel/sgcn.py:154-270) through the oracle's gat_conv."""
    g = H.load(case)
    L, Hd, R, B = [int(v) for v in g["cfg"]]
    b = _batch(H.subjects(g))
    P = H.params(g, "P_gat/", grad=True)
    b["x"].requires_grad_(True)

    def fwd(explain):
        x, w = b["x"], b["edge_attr"]
        if explain:
            m = O.cal_probability(P, x, b["edge_index"], w, R)
            x, w = m["x"], m["w"]
        hs = []
        for l in range(L):
            n = "conv1" if l == 0 else "convs.%d" % (l - 1)
            x = torch.relu(O.gat_conv(x, b["edge_index"], w, P[n + ".lin_src.weight"], P[n + ".att_src"], P[n + ".att_dst"],
                                      P[n + ".lin_edge.weight"], P[n + ".att_edge"], P[n + ".bias"]))
            hs.append(x)
        z = torch.cat(hs, 1).view(B, -1)
        h = torch.relu(z @ P["lin1.weight"].t() + P["lin1.bias"])
        return torch.log_softmax(h @ P["lin2.weight"].t() + P["lin2.bias"], -1)

    o, q = fwd(False), fwd(True)
    lp = O.loss_probability_sgcn(P, b["x"], b["edge_index"], b["edge_attr"], R)
    loss = torch.nn.functional.nll_loss(o, b["y"]) + lp + torch.nn.functional.nll_loss(q, b["y"])
    loss.backward()
    H.assert_close(o, g["gat/logp"])
    H.assert_close(q, g["gat/logp_explain"])
    H.assert_close(loss, g["gat/loss"])
    H.assert_close(b["x"].grad, g["gat/grad/x"], what="grad x")
    for k, v in H.sub_dict(g, "gat/grad/").items():
        if k != "x":
            H.assert_close(P[k].grad, v, what="grad " + k)
