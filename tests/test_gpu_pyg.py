"""The PyG-signature operator seam (ig-gcn_b200/pyg.py) against the oracle: GCNConv / GATConv with the EXACT torch_geometric
call signatures (no csr argument, gradient into edge_weight / edge_attr), to_dense_batch, global pools, scatter -- and an
operator-level restatement of the reference's encoder forward (kernel/sgcn_img_snp.py:207-228) built from them."""
import numpy as np
import pytest
import torch

from oracle import igcn_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _random_graph(N, E, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, N, (E,), generator=g)
    dst = torch.randint(0, N, (E,), generator=g)
    src[:7], dst[:7] = torch.arange(7), torch.arange(7)           # some self loops ...
    src[7], dst[7] = 3, 3                                         # ... one duplicated (the last one wins)
    dst[dst == N - 1] = 0                                         # node N-1 has no in-edge at all
    w = torch.rand(E, generator=g) + 0.1
    return torch.stack([src, dst]), w


@pytest.mark.parametrize("N,E,C,Oc,with_w", [(1000, 5000, 7, 12, True), (37, 90, 3, 16, True), (5000, 20000, 16, 16, False), (64, 0, 4, 4, True)])
def test_gcn_conv_arbitrary_graph(N, E, C, Oc, with_w):
    from igcn_b200 import pyg
    ei, w = _random_graph(N, max(E, 8), 11) if E else (torch.zeros((2, 0), dtype=torch.int64), torch.zeros(0))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, C, generator=g)
    conv = pyg.GCNConv(C, Oc).to(DEV)
    with torch.no_grad():
        conv.bias.uniform_(-0.2, 0.2)
    go = torch.randn(N, Oc, generator=g)
    xc = x.to(DEV).requires_grad_(True)
    wc = w.to(DEV).requires_grad_(True) if with_w else None
    out = conv(xc, ei.to(DEV), wc)                                # the PyG call signature
    (out * go.to(DEV)).sum().backward()
    x64 = x.double().requires_grad_(True)
    w64 = w.double().requires_grad_(True) if with_w else None
    W64, b64 = conv.lin.weight.detach().cpu().double().requires_grad_(True), conv.bias.detach().cpu().double().requires_grad_(True)
    ref = O.gcn_conv(x64, ei, w64, W64, b64)
    (ref * go.double()).sum().backward()
    H.assert_close(out, ref, what="out")
    H.assert_close(xc.grad, x64.grad, what="dx")
    H.assert_close(conv.lin.weight.grad, W64.grad, what="dW")
    H.assert_close(conv.bias.grad, b64.grad, what="db")
    if with_w and E:
        H.assert_close(wc.grad, w64.grad, what="d edge_weight")


def _collated(n, R, seed):
    from igcn_b200 import synthetic as syn
    from igcn_b200.data import Batch, SubjectSet
    sub = syn.make_subjects(n, rois=R, n_snps=8, seed=seed)
    b = Batch.collate(SubjectSet(sub), np.arange(n), torch.device(DEV))
    return sub, b, O.collate(sub, np.arange(n))


def test_reference_encoder_forward_at_operator_level():
    """kernel/sgcn_img_snp.py:207-228 restated with the operators it calls: cal_probability (explain pass: masked features, edge
    weights times the learned edge probability -- so the gradient must flow INTO edge_weight), conv1 / convs with relu, cat,
    to_dense_batch, and the three global pools of the graph_pool branch (:229-236)."""
    from igcn_b200 import pyg
    n, R, Hd, L = 6, 90, 16, 2
    sub, b, c = _collated(n, R, 21)
    torch.manual_seed(0)
    convs = torch.nn.ModuleList([pyg.GCNConv(3, Hd), pyg.GCNConv(Hd, Hd)]).to(DEV)
    prob = (torch.rand(R, 3) - 0.5).to(DEV).requires_grad_(True)
    pb = (torch.rand(6, 1) - 0.5).to(DEV).requires_grad_(True)
    x, ei, w = b.x.clone().requires_grad_(True), b.edge_index, b.edge_attr
    # -- the reference's lines, on the shadowed operators --
    xm = (x.view(n, R, 3) * prob).reshape(n * R, 3)
    pe = torch.sigmoid(torch.cat((xm[ei[0]], xm[ei[1]]), -1) @ pb).view(-1)
    wm = w * pe
    hs, h = [], xm
    for conv in convs:
        h = torch.relu(conv(h, ei, wm))
        hs.append(h)
    hcat = torch.cat(hs, 1)
    dense, mask = pyg.to_dense_batch(hcat, b.batch, fill_value=float(hcat.min().item()) - 1)
    pooled = torch.cat([pyg.global_mean_pool(hcat, b.batch), pyg.global_max_pool(hcat, b.batch), pyg.global_add_pool(hcat, b.batch)], 1)
    gen = torch.Generator().manual_seed(3)
    g1, g2 = torch.randn(dense.shape, generator=gen), torch.randn(pooled.shape, generator=gen)
    ((dense * g1.to(DEV)).sum() + (pooled * g2.to(DEV)).sum()).backward()
    # -- oracle --
    P = {"prob": prob.detach().cpu().double().requires_grad_(True), "prob_bias": pb.detach().cpu().double().requires_grad_(True),
         "conv1.lin.weight": convs[0].lin.weight.detach().cpu().double().requires_grad_(True),
         "conv1.bias": convs[0].bias.detach().cpu().double().requires_grad_(True),
         "convs.0.lin.weight": convs[1].lin.weight.detach().cpu().double().requires_grad_(True),
         "convs.0.bias": convs[1].bias.detach().cpu().double().requires_grad_(True)}
    x64 = torch.from_numpy(c["x"]).double().requires_grad_(True)
    e64 = torch.from_numpy(c["edge_index"])
    m = O.cal_probability(P, x64, e64, torch.from_numpy(c["edge_attr"]).double(), R)
    ref = O.sgcn_encoder(P, m["x"], e64, m["w"], L, R, relu_pattern=(dense.detach().cpu() > 0))
    rp = torch.cat([ref.mean(1), ref.max(1)[0], ref.sum(1)], 1)
    ((ref * g1.double()).sum() + (rp * g2.double()).sum()).backward()
    assert bool(mask.all()) and dense.shape == (n, R, L * Hd)
    H.assert_close(dense, ref, what="to_dense_batch(cat(relu(conv)))")
    H.assert_close(pooled, rp, what="global pools")
    H.assert_close(x.grad, x64.grad, what="dx")
    H.assert_close(prob.grad, P["prob"].grad, what="d prob")
    H.assert_close(pb.grad, P["prob_bias"].grad, what="d prob_bias")
    for i, nme in enumerate(["conv1", "convs.0"]):
        H.assert_close(convs[i].lin.weight.grad, P[nme + ".lin.weight"].grad, what="dW " + nme)
        H.assert_close(convs[i].bias.grad, P[nme + ".bias"].grad, what="db " + nme)


def test_gat_conv_pyg_signature():
    """GATConv(in, out, edge_dim=1)(x, edge_index, edge_attr) as called at kernel/sgcn.py:163-166: no structure argument; the
    gradient reaches edge_attr."""
    from igcn_b200 import pyg
    n, R = 5, 30
    sub, b, c = _collated(n, R, 4)
    torch.manual_seed(1)
    conv = pyg.GATConv(3, 8, edge_dim=1).to(DEV)
    x = b.x.clone().requires_grad_(True)
    ea = b.edge_attr.clone().requires_grad_(True)
    out = conv(x, b.edge_index, ea)
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(2))
    (out * go.to(DEV)).sum().backward()
    d = lambda t: t.detach().cpu().double().requires_grad_(True)
    x64, ea64 = d(b.x), d(b.edge_attr)
    Wp = [d(conv.lin_src.weight), d(conv.att_src), d(conv.att_dst), d(conv.lin_edge.weight), d(conv.att_edge), d(conv.bias)]
    ref = O.gat_conv(x64, torch.from_numpy(c["edge_index"]), ea64, *Wp)
    (ref * go.double()).sum().backward()
    H.assert_close(out, ref, what="gat out")
    H.assert_close(x.grad, x64.grad, what="gat dx")
    H.assert_close(ea.grad, ea64.grad, what="gat d edge_attr")
    for p_, r_ in zip([conv.lin_src.weight, conv.att_src, conv.att_dst, conv.lin_edge.weight, conv.att_edge, conv.bias], Wp):
        H.assert_close(p_.grad, r_.grad.view_as(p_.grad), what="gat param grad")


def test_batch_utilities_general_case():
    """to_dense_batch / global pools / scatter on a RAGGED batch vector (not produced by this package: the general formulation)."""
    from igcn_b200 import pyg
    g = torch.Generator().manual_seed(0)
    sizes = [3, 7, 1, 5]
    batch = torch.cat([torch.full((s,), i) for i, s in enumerate(sizes)]).to(DEV)
    x = torch.randn(sum(sizes), 4, generator=g).to(DEV)
    dense, mask = pyg.to_dense_batch(x, batch, fill_value=-9.0)
    assert dense.shape == (4, 7, 4) and mask.sum().item() == sum(sizes)
    off = 0
    for i, s in enumerate(sizes):
        assert torch.equal(dense[i, :s], x[off:off + s]) and bool((dense[i, s:] == -9.0).all()) and bool(mask[i, :s].all()) and not bool(mask[i, s:].any())
        H.assert_close(pyg.global_add_pool(x, batch)[i], x[off:off + s].sum(0), what="add pool")
        H.assert_close(pyg.global_mean_pool(x, batch)[i], x[off:off + s].mean(0), what="mean pool")
        assert torch.equal(pyg.global_max_pool(x, batch)[i], x[off:off + s].max(0)[0])
        off += s
    # torch_scatter.scatter as kernel/go_model.py:200 calls it: (B, nnz, d) summed over dim 1 into rows
    src = torch.randn(3, 9, 5, generator=g).to(DEV).requires_grad_(True)
    index = torch.tensor([0, 0, 1, 3, 3, 3, 4, 4, 6], device=DEV)
    out = pyg.scatter(src, index, dim=1, reduce="sum")
    ref = torch.zeros(3, 7, 5, device=DEV).index_add(1, index, src.detach())
    assert out.shape == (3, 7, 5)
    H.assert_close(out, ref, what="scatter sum")
    out.sum().backward()
    assert torch.equal(src.grad, torch.ones_like(src))


def test_graph_csr_cache_is_not_fooled_by_a_recycled_address():
    """A loader with fixed-size batches frees one edge_index and allocates the next one, same shape, very likely at the same
    device address: the structure cache must not return the previous batch's CSR for it."""
    from igcn_b200 import pyg
    N, E, C = 200, 900, 5
    conv = pyg.GCNConv(C, 8).to(DEV)
    x = torch.randn(N, C, generator=torch.Generator().manual_seed(0))
    xc = x.to(DEV)
    seen = set()
    for seed in (21, 22, 23, 24):
        ei, w = _random_graph(N, E, seed)
        ei_dev = ei.to(DEV)
        seen.add(ei_dev.data_ptr())
        out = conv(xc, ei_dev, w.to(DEV))
        ref = O.gcn_conv(x.double(), ei, w.double(), conv.lin.weight.detach().cpu().double(), conv.bias.detach().cpu().double())
        H.assert_close(out, ref, what="out seed %d" % seed)
        del ei_dev, out
    # in-place edit of a cached edge_index: the version counter changes the key
    ei, w = _random_graph(N, E, 30)
    ei_dev = ei.to(DEV)
    conv(xc, ei_dev, w.to(DEV))
    ei2, _ = _random_graph(N, E, 31)
    ei_dev.copy_(ei2.to(DEV))
    out = conv(xc, ei_dev, w.to(DEV))
    ref = O.gcn_conv(x.double(), ei2, w.double(), conv.lin.weight.detach().cpu().double(), conv.bias.detach().cpu().double())
    H.assert_close(out, ref, what="out after in-place edit")
