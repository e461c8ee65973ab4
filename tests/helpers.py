"""Shared test plumbing: golden-fixture loading and the ONE statement of the parity tolerances.

Tolerances (north star: "within 1e-4 relative of the reference PyG implementation"):

  rule A  `assert_close(got, ref)`:  every element satisfies |got - ref| <= RTOL * (|ref| + rms(ref)) with RTOL = 1e-4.
          Element-relative with a floor of one part in 1e4 of the tensor's own RMS -- an fp32 result cannot be relatively
          accurate on elements that are cancellation residues many orders below the tensor's scale, and the reference's own fp32
          output is not either.  Integer / index tensors are compared with array_equal by the tests, never through here.

  rule B  `assert_parity(got, ref32, truth64)`:  rule A against the reference's fp32 result, OR -- SURVEY.md 8(c)'s second
          criterion, for long gradient chains where two correct fp32 evaluations in different summation orders differ by more
          than 1e-4 -- the CUDA result is not further from the fp64 truth than the fp32 reference itself is, by more than 2x:
              err(got, truth64) <= 2 * err(ref32, truth64)
          where err is the rule-A error measure.  Which rule admitted each tensor is recorded in PARITY_LOG and written to
          gpurun_out/parity_report.json at the end of the session (tests/conftest.py).
"""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

RTOL = 1e-4
PARITY_LOG = []          # dicts: what, rule, err32, err_got64, err_ref64


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def sub_dict(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def seeded_param(seed, bound, shape):
    """Regenerates a parameter stored as a recipe by tests/golden/make_golden.py::sd_np(seeded=True)."""
    gen = torch.Generator().manual_seed(int(seed))
    return (torch.rand(tuple(int(s) for s in shape), generator=gen, dtype=torch.float32) * 2.0 - 1.0) * float(bound)


def state_arrays(g, prefix="P/"):
    """name -> torch tensor for every stored parameter / buffer under `prefix`, including the seeded recipes."""
    out = {k: torch.from_numpy(v) for k, v in sub_dict(g, prefix).items()}
    for k, v in sub_dict(g, prefix.rstrip("/") + "seed/").items():
        out[k] = seeded_param(v[0], v[1], v[2:])
    return out


def params(g, prefix="P/", dtype=torch.float32, device="cpu", grad=False):
    out = {}
    for k, t in state_arrays(g, prefix).items():
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


def _np(a):
    return np.asarray(a.detach().cpu() if torch.is_tensor(a) else a, dtype=np.float64)


def rel_err(a, b):
    """max over elements of |a-b| / (|b| + rms(b)): 1.0 means 'off by the element's own size plus the tensor's RMS'."""
    a, b = _np(a), _np(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    rms = float(np.sqrt(np.mean(b * b)))
    den = np.abs(b) + max(rms, 1e-30)
    return float((np.abs(a - b) / den).max())


def assert_close(a, b, rtol=RTOL, what=""):
    e = rel_err(a, b)
    assert e <= rtol, "%s: rel err %.3e > %.1e" % (what, e, rtol)
    return e


class Collector(object):
    """Collects parity failures so one GPU run reports every offending tensor, then fails once: `with Collector() as c:
    c.parity(...)`."""

    def __init__(self):
        self.failures = []

    def __enter__(self):
        return self

    def parity(self, *a, **kw):
        try:
            return assert_parity(*a, **kw)
        except AssertionError as e:
            self.failures.append(str(e))

    def close(self, *a, **kw):
        try:
            return assert_close(*a, **kw)
        except AssertionError as e:
            self.failures.append(str(e))

    def __exit__(self, et, ev, tb):
        if et is None and self.failures:
            raise AssertionError("%d tensors out of tolerance:\n  " % len(self.failures) + "\n  ".join(self.failures))
        return False


def assert_parity(got, ref32, truth64, what="", rtol=RTOL):
    """Rule A against ref32, else rule B against the fp64 truth (see the module docstring).  Returns the rule that applied."""
    e32 = rel_err(got, ref32)
    rec = dict(what=what, err32=e32)
    if e32 <= rtol:
        rec["rule"] = "A"
        PARITY_LOG.append(rec)
        return "A"
    assert truth64 is not None, "%s: rel err %.3e > %.1e and no fp64 truth to apply rule B" % (what, e32, rtol)
    eg, er = rel_err(got, truth64), rel_err(ref32, truth64)
    rec.update(rule="B", err_got64=eg, err_ref64=er)
    PARITY_LOG.append(rec)
    assert eg <= 2.0 * er, ("%s: %.3e from the fp32 reference (> %.0e) and %.3e from the fp64 truth, more than twice the fp32 "
                            "reference's own %.3e" % (what, e32, rtol, eg, er))
    return "B"
