"""On-device input pipeline (ig-gcn_b200/device_data.py, csrc/gdc.cu): graph diffusion convolution + COO emission on the GPU against
the numpy restatement of util_gdc.py (synthetic.gdc_topk / make_subjects) -- integer structure BIT EXACT for the same
connectivity matrices -- and device-side collation against the host path."""
import numpy as np
import pytest
import torch

from oracle import igcn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n,R,k", [(24, 90, 3), (6, 264, 3), (10, 33, 5), (3, 7, 2)])
def test_gdc_on_device_matches_numpy_restatement(n, R, k):
    from igcn_b200 import synthetic as syn
    from igcn_b200.device_data import gdc_topk_device
    rngs = [np.random.default_rng([77, i]) for i in range(n)]
    conn = syn.knn_connectivity(rngs, R, knn=min(5, R - 1))                      # (n, R, R) fp64, symmetric
    diff = syn.gdc_topk(conn, k=k)                                               # util_gdc.py:7-14 + 25-31 on the host
    src, dst, w = gdc_topk_device(torch.from_numpy(conn).to(DEV), k=k)
    src, dst, w = src.cpu().numpy(), dst.cpu().numpy(), w.cpu().numpy()
    for b in range(n):
        a32 = diff[b].astype(np.float32)
        s, d = np.nonzero(a32)                                                   # row-major == scipy coo_matrix(dense) order
        assert s.size == R * k
        assert np.array_equal(src[b], s.astype(np.int32)) and np.array_equal(dst[b], d.astype(np.int32)), b
        assert np.abs(w[b] - a32[s, d]).max() <= 2e-7, b
        assert np.array_equal(np.bincount(dst[b], minlength=R), np.full(R, k))    # exactly k in-edges per node


def test_device_subject_set_and_collation():
    from igcn_b200.data import Batch, SubjectSet
    from igcn_b200.device_data import DeviceSubjectSet, collate_device
    dev = torch.device(DEV)
    ds = DeviceSubjectSet.generate(70, rois=90, n_snps=54, seed=3, device=dev, chunk=32)
    assert ds.x.shape == (70, 90, 3) and ds.edge_src.shape == (70, 270) and ds.snps_feat.shape == (70, 54) and len(ds) == 70
    assert float(ds.edge_attr.min()) > 0 and float(ds.x.min()) >= 0 and float(ds.x.max()) < 1
    # column-normalised diffusion weights: the in-edges of every node sum to 1 (util_gdc.py:29-30)
    insum = torch.zeros(70, 90, device=dev).scatter_add_(1, ds.edge_dst.long(), ds.edge_attr)
    assert float((insum - 1).abs().max()) < 1e-5
    idx = torch.tensor([5, 0, 69, 33, 12, 12, 7], device=dev)
    b = collate_device(ds, idx)
    # the same subjects through the HOST path give the same batch, bit for bit
    packed = dict(x=ds.x.cpu().numpy(), edge_ptr=np.arange(71, dtype=np.int64) * 270, edge_src=ds.edge_src.cpu().numpy().reshape(-1),
                  edge_dst=ds.edge_dst.cpu().numpy().reshape(-1), edge_attr=ds.edge_attr.cpu().numpy().reshape(-1),
                  snps_feat=ds.snps_feat.cpu().numpy(), y=ds.y.cpu().numpy(), clini_score=ds.clini_score.cpu().numpy(),
                  tsne_fdim=ds.tsne_fdim.cpu().numpy(), clust_y=ds.clust_y.cpu().numpy(), sbjID=ds.sbjID.cpu().numpy())
    h = Batch.collate(SubjectSet(packed), idx.cpu().numpy(), dev)
    for name in ("x", "edge_index", "edge_attr", "batch", "snps_feat", "y", "clini_score", "tsne_fdim", "clust_y", "sbjID"):
        assert torch.equal(getattr(b, name), getattr(h, name)), name
    for name in ("rowptr_t", "csr_src", "csr_perm", "csr_w", "rowptr_s", "csc_pos"):
        assert torch.equal(getattr(b.csr, name), getattr(h.csr, name)), name
    c = O.collate(packed, idx.cpu().numpy())
    assert np.array_equal(b.edge_index.cpu().numpy(), c["edge_index"])
    # in-place re-collation into the static buffers of a captured step
    b2 = collate_device(ds, torch.tensor([1, 2, 3, 4, 5, 6, 8], device=dev), out=b)
    assert b2 is b and int(b.sbjID[0]) == 1
