import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # belt and braces: a gpu-marked test never runs without a device
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if not has:
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Which tolerance rule admitted each tensor (tests/helpers.py) -> gpurun_out/parity_report.json."""
    try:
        import json
        from tests import helpers as H
        if H.PARITY_LOG:
            d = os.path.join(ROOT, "gpurun_out")
            os.makedirs(d, exist_ok=True)
            notes = [r for r in H.PARITY_LOG if "ReLU sign flips" in r["what"]]
            H.PARITY_LOG[:] = [r for r in H.PARITY_LOG if "ReLU sign flips" not in r["what"]]
            nb = sum(1 for r in H.PARITY_LOG if r["rule"] == "B")
            with open(os.path.join(d, "parity_report.json"), "w") as f:
                json.dump(dict(rtol=H.RTOL, tensors=len(H.PARITY_LOG), rule_B=nb, relu_sign_flips=[r["what"] for r in notes],
                               worst_rule_A=max([r["err32"] for r in H.PARITY_LOG if r["rule"] == "A"] or [0.0]),
                               rule_B_tensors=[r for r in H.PARITY_LOG if r["rule"] == "B"],
                               worst_rule_A_tensors=sorted([r for r in H.PARITY_LOG if r["rule"] == "A"],
                                                           key=lambda r: -r["err32"])[:12]), f, indent=1)
    except Exception:
        pass
