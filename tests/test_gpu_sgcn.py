"""GPU parity: device collation (bit exact) and the fused SGCN encoder fwd/bwd vs the oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import igcn_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def _subjects(n, R, seed, ragged=False, S=12):
    from igcn_b200 import synthetic as syn
    sub = syn.make_subjects(n, rois=R, n_snps=S, seed=seed)
    if ragged:
        rng = np.random.default_rng(seed)
        E = sub["edge_src"].size
        keep = rng.random(E) > 0.15
        ep = sub["edge_ptr"]
        # graph 1 loses ALL its self loops; graph 0 gets a duplicated self loop on node 0 (last one must win)
        g1 = np.arange(ep[1], ep[2])
        keep[g1[sub["edge_src"][g1] == sub["edge_dst"][g1]]] = False
        cnt = np.array([keep[ep[i]:ep[i + 1]].sum() for i in range(n)])
        for k in ("edge_src", "edge_dst", "edge_attr"):
            sub[k] = sub[k][keep]
        sub["edge_ptr"] = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        # append a duplicate self loop (0,0) with another weight to graph 0
        e1 = int(sub["edge_ptr"][1])
        for k, v in (("edge_src", 0), ("edge_dst", 0), ("edge_attr", 0.37)):
            sub[k] = np.insert(sub[k], e1, v).astype(sub[k].dtype)
        sub["edge_ptr"][1:] += 1
    return sub


def test_collate_golden_bit_exact():
    from igcn_b200.data import Batch, SubjectSet
    g = H.load("collate_r30")
    ss = SubjectSet(H.subjects(g))
    b = Batch.collate(ss, g["idx"], _dev())
    torch.cuda.synchronize()
    for k in ("x", "edge_index", "edge_attr", "batch", "snps_feat", "y", "clini_score", "tsne_fdim", "clust_y", "sbjID"):
        got = getattr(b, k).cpu().numpy()
        assert got.dtype == g["out/" + k].dtype, k
        assert np.array_equal(got, g["out/" + k]), k
    assert b.num_graphs == int(g["out/num_graphs"])


@pytest.mark.parametrize("n,R,ragged", [(5, 30, True), (16, 90, False), (3, 264, True), (300, 7, True)])
def test_collate_csr_matches_stable_sort(n, R, ragged):
    from igcn_b200.data import Batch, SubjectSet
    sub = _subjects(n, R, seed=n + R, ragged=ragged)
    idx = np.random.default_rng(0).permutation(n)
    c = O.collate(sub, idx)
    b = Batch.collate(SubjectSet(sub), idx, _dev())
    torch.cuda.synchronize()
    assert np.array_equal(b.edge_index.cpu().numpy(), c["edge_index"])
    assert np.array_equal(b.batch.cpu().numpy(), c["batch"])
    rowptr, src, perm = O.target_sorted_csr(c["edge_index"], n * R)
    colptr, pos = O.source_sorted_csr(c["edge_index"], n * R)
    csr = b.csr
    assert np.array_equal(csr.rowptr_t.cpu().numpy(), rowptr)
    assert np.array_equal(csr.csr_src.cpu().numpy(), src)
    assert np.array_equal(csr.csr_perm.cpu().numpy(), perm)
    assert np.array_equal(csr.csr_w.cpu().numpy(), c["edge_attr"][perm])
    assert np.array_equal(csr.rowptr_s.cpu().numpy(), colptr)
    assert np.array_equal(csr.csc_pos.cpu().numpy(), pos)
    # the same structure from an already collated int64 edge_index
    b2 = Batch.from_device_tensors(b.x, b.edge_index, b.edge_attr, R)
    for k in ("rowptr_t", "csr_src", "csr_perm", "csr_w", "rowptr_s", "csc_pos"):
        assert torch.equal(getattr(b2.csr, k), getattr(csr, k)), k


def test_collate_empty_batch():
    from igcn_b200.data import Batch, SubjectSet
    sub = _subjects(2, 10, seed=1)
    b = Batch.collate(SubjectSet(sub), np.zeros(0, np.int64), _dev())
    assert b.num_graphs == 0 and b.edge_index.shape == (2, 0) and b.batch.numel() == 0


def _enc_params(L, Hd, R, F0, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    P = {"prob": (torch.rand(R, F0, generator=g) * 2 - 1) / np.sqrt(F0), "prob_bias": torch.rand(2 * F0, 1, generator=g) * 2 - 1}
    for l in range(L):
        n = "conv1" if l == 0 else "convs.%d" % (l - 1)
        fin = F0 if l == 0 else Hd
        P[n + ".lin.weight"] = (torch.rand(Hd, fin, generator=g) * 2 - 1) * np.sqrt(6.0 / (fin + Hd))
        P[n + ".bias"] = (torch.rand(Hd, generator=g) * 2 - 1) * 0.1
    return {k: v.to(dtype) for k, v in P.items()}


@pytest.mark.parametrize("n,R,L,Hd,ragged", [(6, 30, 3, 8, True), (16, 90, 2, 16, False), (5, 264, 2, 16, True),
                                               (9, 90, 4, 5, False), (7, 90, 2, 10, True), (4, 90, 1, 16, False),
                                               (11, 33, 2, 16, True), (3, 288, 2, 16, False), (40, 7, 2, 16, True),
                                               (333, 90, 2, 16, False), (301, 90, 2, 16, True)])   # > 148 graphs: several graphs per CTA
@pytest.mark.parametrize("explain", [False, True])
def test_sgcn_encoder_fwd_bwd(n, R, L, Hd, ragged, explain):
    from igcn_b200 import ops
    from igcn_b200.data import Batch, SubjectSet
    dev = _dev()
    sub = _subjects(n, R, seed=3 * n + R + L, ragged=ragged)
    idx = np.arange(n)
    c = O.collate(sub, idx)
    b = Batch.collate(SubjectSet(sub), idx, dev)
    P64 = _enc_params(L, Hd, R, 3, seed=L * 100 + Hd, dtype=torch.float64)
    for v in P64.values():
        v.requires_grad_(True)
    ei = torch.from_numpy(c["edge_index"])
    gen = torch.Generator().manual_seed(5)
    g_out = torch.randn((n, R, L * Hd), generator=gen, dtype=torch.float64)
    g_pe = torch.randn(ei.shape[1], generator=gen, dtype=torch.float64) if explain else None
    # CUDA path -------------------------------------------------------------------------------
    Pc = {k: v.detach().float().to(dev).requires_grad_(True) for k, v in P64.items()}
    xc = b.x.clone().requires_grad_(True)
    Ws = [Pc[("conv1" if l == 0 else "convs.%d" % (l - 1)) + ".lin.weight"] for l in range(L)]
    bs = [Pc[("conv1" if l == 0 else "convs.%d" % (l - 1)) + ".bias"] for l in range(L)]
    out, pe = ops.sgcn_encoder(xc, b.csr, Ws, bs, Pc["prob"] if explain else None, Pc["prob_bias"] if explain else None,
                               want_pe=explain)
    perm = b.csr.csr_perm.long()
    lossc = (out * g_out.float().to(dev)).sum()
    if explain:
        lossc = lossc + (pe * g_pe.float().to(dev)[perm]).sum()
    lossc.backward()
    torch.cuda.synchronize()
    # oracle in fp64 (truth): values with its own ReLU; gradients with the KERNEL's active set imposed (oracle._relu: an fp32
    # kernel and the fp64 oracle may disagree on the sign of a pre-activation within rounding distance of zero -- about one
    # element per million -- and every such flip moves the gradients through it by a full term) -------------------------------
    x64 = torch.from_numpy(c["x"]).double().requires_grad_(True)
    w64 = torch.from_numpy(c["edge_attr"]).double()
    pattern = out.detach().cpu() > 0
    pre = []
    if explain:
        m = O.cal_probability(P64, x64, ei, w64, R)
        ref = O.sgcn_encoder(P64, m["x"], ei, m["w"], L, R, relu_pattern=pattern, pre_out=pre)
        pe_ref = m["p_e"]
    else:
        ref = O.sgcn_encoder(P64, x64, ei, w64, L, R, relu_pattern=pattern, pre_out=pre)
        pe_ref = None
    prez = torch.cat(pre, 1).view(n, R, L * Hd)
    flips = pattern != (prez > 0)
    assert int(flips.sum()) <= max(2, int(1e-5 * flips.numel())), "%d ReLU sign flips" % int(flips.sum())
    if bool(flips.any()):
        assert float(prez[flips].abs().max()) <= 2e-5 * float(prez.abs().max()), "a flipped pre-activation is not near zero"
        H.PARITY_LOG.append(dict(what="encoder n=%d R=%d explain=%s: ReLU sign flips vs fp64" % (n, R, explain), rule="A",
                                 err32=float(flips.sum())))
    loss = (ref * g_out).sum() + ((pe_ref * g_pe).sum() if explain else 0.0)
    loss.backward()
    H.assert_close(out, torch.relu(prez), what="encoder out")
    if explain:
        H.assert_close(pe, pe_ref[perm.cpu()], what="p_e")
    H.assert_close(xc.grad, x64.grad, what="dx")
    names = [k for k in Pc if explain or not k.startswith("prob")]
    for k in names:
        H.assert_close(Pc[k].grad, P64[k].grad, what="grad " + k)
    if not explain:
        assert Pc["prob"].grad is None


def test_register_tiled_kernels_on_default_shape(monkeypatch):
    """F0=3/H=16/L=2 takes the tensor-core kernels (sgcn_mma.cuh); IGCN_SGCN_MMA=0 sends the same shape through the register-tiled
    kernels of sgcn_fast.cuh (the path of L != 2), so both stay covered."""
    monkeypatch.setenv("IGCN_SGCN_MMA", "0")
    test_sgcn_encoder_fwd_bwd(6, 90, 2, 16, True, True)
    test_sgcn_encoder_fwd_bwd(5, 264, 2, 16, True, False)


def test_generic_kernels_on_default_shape(monkeypatch):
    """F0=3/H=16 normally takes the register-tiled kernels; IGCN_FORCE_GENERIC=1 sends the same shape through the
    shape-generic kernels so both stay covered."""
    monkeypatch.setenv("IGCN_FORCE_GENERIC", "1")
    test_sgcn_encoder_fwd_bwd(6, 90, 2, 16, True, True)
    test_sgcn_encoder_fwd_bwd(6, 90, 2, 16, False, False)


def test_sgcn_encoder_deterministic_and_full_size_properties():
    """BASELINE config-4 shape (R=264, B=4096 scaled to what generates in seconds: B=512): run-to-run bit
    identical (no float atomics), and the plain pass is linear in x before the first ReLU => homogeneity
    relu(conv(a*x)) = a*relu(conv(x)) for a>0 when biases are zero."""
    from igcn_b200 import ops
    from igcn_b200.data import Batch, SubjectSet
    dev = _dev()
    n, R, L, Hd = 512, 264, 2, 16
    sub = _subjects(n, R, seed=99)
    b = Batch.collate(SubjectSet(sub), np.arange(n), dev)
    P = _enc_params(L, Hd, R, 3, seed=1, dtype=torch.float32)
    Ws = [P[("conv1" if l == 0 else "convs.%d" % (l - 1)) + ".lin.weight"].to(dev).requires_grad_(True) for l in range(L)]
    bs = [torch.zeros(Hd, device=dev, requires_grad=True) for _ in range(L)]
    prob, pb = P["prob"].to(dev).requires_grad_(True), P["prob_bias"].to(dev).requires_grad_(True)
    outs, grads = [], []
    for _ in range(2):
        x = b.x.clone().requires_grad_(True)
        o, pe = ops.sgcn_encoder(x, b.csr, Ws, bs, prob, pb, want_pe=True)
        (o.sum() + pe.sum()).backward()
        outs.append(o.detach().clone())
        grads.append([x.grad.clone(), prob.grad.clone(), Ws[0].grad.clone(), Ws[1].grad.clone()])
        prob.grad = None
        Ws[0].grad = None
        Ws[1].grad = None
    assert torch.equal(outs[0], outs[1])
    for a, c in zip(*grads):
        assert torch.equal(a, c)
    # ... and the values themselves against the oracle at this size: fp32 oracle = reference, fp64 oracle = truth (rule A / B)
    c = O.collate(sub, np.arange(n))
    ei = torch.from_numpy(c["edge_index"])
    res = {}
    for dt in (torch.float32, torch.float64):
        Po = {k: v.to(dt).clone().requires_grad_(True) for k, v in P.items()}
        for k in list(Po):
            if k.endswith(".bias"):
                Po[k] = torch.zeros_like(Po[k]).requires_grad_(True)
        xo = torch.from_numpy(c["x"]).to(dt).requires_grad_(True)
        mm = O.cal_probability(Po, xo, ei, torch.from_numpy(c["edge_attr"]).to(dt), R)
        ref = O.sgcn_encoder(Po, mm["x"], ei, mm["w"], L, R)
        (ref.sum() + mm["p_e"].sum()).backward()
        res[dt] = (ref.detach(), xo.grad, Po["prob"].grad, Po["conv1.lin.weight"].grad, Po["convs.0.lin.weight"].grad)
    got = [outs[0]] + grads[0]
    for nm, a, r32, r64 in zip(["out", "dx", "d prob", "dW1", "dW2"], got, res[torch.float32], res[torch.float64]):
        H.assert_parity(a, r32, r64, what="encoder n=512 R=264 " + nm)
    o1, _ = ops.sgcn_encoder(b.x, b.csr, Ws, bs)
    o2, _ = ops.sgcn_encoder(b.x * 3.0, b.csr, Ws, bs)
    H.assert_close(o2, o1 * 3.0, what="homogeneity")
    assert o1.shape == (n, R, L * Hd)


def test_dropout_mask_bank_statistics_and_replay():
    """One-launch Philox masks: values are 0 or 1/keep, keep rates match, consecutive passes differ, same seed replays."""
    from igcn_b200.ops import MaskBank
    dev = _dev()

    def run(bank, n_pass):
        outs = []
        for _ in range(n_pass):
            bank.begin_pass(64, dev)
            a = bank.get("a", (64, 54, 1), 0.4).clone()
            b = bank.get("b", (64, 301), 0.5).clone()
            bank.end_pass()
            outs.append((a, b))
        return outs
    o = run(MaskBank(seed=123), 4)            # pass 0 records with torch, passes 1..3 come from the kernel
    for a, b in o[1:]:
        assert bool(((a == 0) | ((a - 1 / 0.6).abs() < 1e-6)).all())
        assert bool(((b == 0) | (b == 2.0)).all())
        assert abs((a > 0).float().mean().item() - 0.6) < 0.03
        assert abs((b > 0).float().mean().item() - 0.5) < 0.02
    assert not torch.equal(o[1][0], o[2][0]) and not torch.equal(o[2][1], o[3][1])
    o2 = run(MaskBank(seed=123), 4)
    assert torch.equal(o[2][0], o2[2][0]) and torch.equal(o[3][1], o2[3][1])


def test_dropout_masks_of_consecutive_calls_are_uncorrelated():
    """Consecutive launches must use disjoint Philox counter blocks: the mask of call c+1 is not a shifted copy of call c
    (the offset counts 32-bit outputs and every thread consumes four per call)."""
    from igcn_b200.ops import MaskBank
    dev = _dev()
    bank = MaskBank(seed=7)
    outs = []
    for _ in range(6):
        bank.begin_pass(0, dev)
        outs.append(bank.get("m", (1 << 16,), 0.5).clone())
        bank.end_pass()
    for a, b in zip(outs[1:-1], outs[2:]):          # outs[0] is the torch-drawn recording pass
        ka, kb = (a > 0).float(), (b > 0).float()
        for shift in range(0, 4):
            n = ka.numel() - shift
            agree = (ka[shift:shift + n] == kb[:n]).float().mean().item()
            agree2 = (kb[shift:shift + n] == ka[:n]).float().mean().item()
            assert abs(agree - 0.5) < 0.02 and abs(agree2 - 0.5) < 0.02, (shift, agree, agree2)
