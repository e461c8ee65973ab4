"""Shared test plumbing: golden-fixture loading and tolerant comparison."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north star: "within 1e-4 relative" for floating point; integer/index work bit exact.
RTOL = 1e-4


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def sub_dict(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def params(g, prefix="P/", dtype=torch.float32, device="cpu", grad=False):
    out = {}
    for k, v in sub_dict(g, prefix).items():
        t = torch.from_numpy(v)
        if t.is_floating_point():
            t = t.to(dtype)
        t = t.to(device)
        if grad and t.is_floating_point() and "running" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


def subjects(g, prefix="sub/"):
    d = sub_dict(g, prefix)
    for k in ("rois", "n_snps"):
        if k in d:
            d[k] = int(d[k])
    return d


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny): the 'relative to the tensor's scale' error used for parity."""
    a = np.asarray(a.detach().cpu() if torch.is_tensor(a) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if torch.is_tensor(b) else b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    scale = max(np.abs(b).max(), 1e-12)
    return float(np.abs(a - b).max() / scale)


def assert_close(a, b, rtol=RTOL, what=""):
    e = rel_err(a, b)
    assert e <= rtol, "%s: rel err %.3e > %.1e" % (what, e, rtol)
