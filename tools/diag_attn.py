#!/usr/bin/env python
"""Debug helper: runs the cross attention fwd+bwd on one shape and dumps / compares the saved tables and every gradient.
    IGCN_ATTN_SCALAR_TOKENS=1 python tools/diag_attn.py dump scalar ; python tools/diag_attn.py dump mma ; python tools/diag_attn.py cmp scalar mma"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

B, R, M, E = 5, 90, 19, 32


def dump(tag):
    import __graft_entry__ as ge
    ge.build()
    from igcn_b200 import ops
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    mha = torch.nn.MultiheadAttention(E, 2, batch_first=True).to(dev)
    with torch.no_grad():
        mha.in_proj_bias.uniform_(-0.2, 0.2)
        mha.out_proj.bias.uniform_(-0.2, 0.2)
    q = torch.randn(B, R, E, device=dev, requires_grad=True)
    kv = torch.randn(B, M, E, device=dev, requires_grad=True)
    g = torch.randn(B, R, E, device=dev)
    res = {}
    for rep in range(3):
        q.grad = kv.grad = None
        mha.zero_grad()
        out = ops.cross_attention(q, kv, mha, relu=True)
        tab = [t for t in out.grad_fn.saved_tensors if t.dim() == 2 and t.shape[0] == B and t.shape[1] > 1000]
        (out * g).sum().backward()
        torch.cuda.synchronize()
        res[rep] = dict(out=out.detach().cpu(), tab=tab[0].cpu() if tab else None, dq=q.grad.cpu().clone(), dkv=kv.grad.cpu().clone(),
                        **{"g_" + k: p.grad.cpu().clone() for k, p in mha.named_parameters()})
    # fp64 truth
    ref = torch.nn.MultiheadAttention(E, 2, batch_first=True).to(dev).double()
    ref.load_state_dict({k: v.double() for k, v in mha.state_dict().items()})
    q64, kv64 = q.detach().double().requires_grad_(True), kv.detach().double().requires_grad_(True)
    o64 = torch.relu(ref(q64, kv64, kv64, need_weights=False)[0])
    (o64 * g.double()).sum().backward()
    res["truth"] = dict(out=o64.detach().cpu(), dq=q64.grad.cpu(), dkv=kv64.grad.cpu(), **{"g_" + k: p.grad.cpu() for k, p in ref.named_parameters()})
    torch.save(res, os.path.join(ROOT, "gpurun_out", "diag_attn_%s.pt" % tag))
    for rep in range(3):
        for k, v in res[rep].items():
            if k in res["truth"]:
                d = (v.double() - res["truth"][k]).abs()
                print(tag, "rep", rep, k, "max abs err %.3e (scale %.3e)" % (float(d.max()), float(res["truth"][k].abs().max())),
                      "worst idx", tuple(int(i) for i in torch.nonzero(d == d.max())[0]))
    for k in res[0]:
        if res[0][k] is not None:
            print(tag, k, "run-to-run identical:", all(torch.equal(res[0][k], res[r][k]) for r in (1, 2)))


def cmp(a, b):
    ra = torch.load(os.path.join(ROOT, "gpurun_out", "diag_attn_%s.pt" % a))
    rb = torch.load(os.path.join(ROOT, "gpurun_out", "diag_attn_%s.pt" % b))
    ta, tb = ra[0]["tab"], rb[0]["tab"]
    if ta is not None and tb is not None:
        MP, TS, H = 24, 36, 2
        sec = dict(Kp=(0, H * MP * TS), Vp=(H * MP * TS, 2 * H * MP * TS), c=(2 * H * MP * TS, 2 * H * MP * TS + H * MP),
                   K=(2 * H * MP * TS + H * MP, 2 * H * MP * TS + H * MP + M * 32), V=(2 * H * MP * TS + H * MP + M * 32, ta.shape[1]))
        for n, (lo, hi) in sec.items():
            x, y = ta[:, lo:hi], tb[:, lo:hi]
            if n in ("Kp", "Vp"):
                x, y = x.reshape(B, H, MP, TS)[..., :32], y.reshape(B, H, MP, TS)[..., :32]
            d = (x - y).abs()
            print("tab section", n, "max abs diff %.3e" % float(d.max()), "scale %.3e" % float(x.abs().max()))
    for k in ("out", "dq", "dkv"):
        d = (ra[0][k] - rb[0][k]).abs()
        print(k, "max abs diff %.3e" % float(d.max()), "rows with diff > 1e-5:", sorted(set((int(i[0]), int(i[1])) for i in torch.nonzero(d > 1e-5)))[:40])


if __name__ == "__main__":
    if sys.argv[1] == "dump":
        dump(sys.argv[2])
    else:
        cmp(sys.argv[2], sys.argv[3])
