"""ORACLE / TEST INFRASTRUCTURE ONLY -- not product code.

Pure-torch CPU restatement of the slice of `torch_geometric==2.0.2`
(pinned at /root/reference/environment.yml:183) that the IG-GCN hot path calls.
The real wheel is not installed in this image and cannot be fetched, so the
published 2.0.2 semantics are restated here.

Used for two things only:
  * `tests/golden/make_golden.py` imports the UNMODIFIED reference model files
    from /root/reference on top of this shim to generate golden vectors;
  * nothing in the product (`ig-gcn_b200/`) imports it.
"""
from . import nn, utils, data  # noqa: F401


def is_debug_enabled():
    # reference call site: batch.py:120
    return False
