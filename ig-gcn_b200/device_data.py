"""On-device input pipeline: synthetic ADNI-shaped subjects are GENERATED and PREPROCESSED on the
GPU and minibatches are collated from device-resident arrays -- nothing is staged on the host.  This is what BASELINE configs[4]
needs: 1 M subjects are 3.2 GB of features and ~9.5 GB of diffusion edges, generated per rank.

Reference pipeline being restated (host side there): connectivity matrix -> graph diffusion convolution
(util_gdc.py:7-14 PPR with alpha = 0.05; :25-31 top-k per column + column normalisation; :84-101 dense -> COO), applied per subject
as a `pre_transform` (sgcn_data.py:332-338, main.py:45,193), then Data objects (sgcn_data.py:257-288).

  * the R x R PPR inverse is a batched fp64 inverse (torch.linalg.inv on the device: a LAPACK-style library call, as np.linalg.inv
    is in the reference);
  * the irregular part -- per-column top-k, normalisation, row-major COO emission -- is `igcn_gdc_topk_emit` (csrc/gdc.cu); for the
    same PPR matrix the edge list is bit-identical to the numpy restatement in synthetic.py (tests/test_gpu_device_data.py);
  * `DeviceSubjectSet` has the SubjectSet fields as device tensors; `collate_device` gathers a minibatch with device index_selects
    and runs the same `igcn_collate_csr` kernel as the host path (Batch._finish).

Random inputs use torch's device generator (Philox), so they are NOT the numpy streams of synthetic.make_subjects: same
distributions and shapes, different samples.
"""
from __future__ import annotations

import torch

from . import _lib
from .data import Batch


def gdc_topk_device(adj: torch.Tensor, alpha: float = 0.05, k: int = 3):
    """adj (B,R,R) symmetric, non-negative, positive row sums (any float dtype, CUDA).  Returns LOCAL (edge_src, edge_dst) int32
    (B, R*k) and edge_attr f32 (B, R*k) in row-major COO order per subject -- synthetic.gdc_topk + the COO emission of
    synthetic.make_subjects, on the device."""
    _lib.require_cuda(adj)
    B, R, _ = adj.shape
    a = adj.double()
    dinv = a.sum(-1).rsqrt()
    H = dinv[:, :, None] * a * dinv[:, None, :]
    eye = torch.eye(R, dtype=torch.float64, device=adj.device)
    ppr = alpha * torch.linalg.inv(eye[None] - (1.0 - alpha) * H)          # util_gdc.py:7-14
    ppr = ppr.contiguous()
    src = torch.empty((B, R * k), dtype=torch.int32, device=adj.device)
    dst = torch.empty_like(src)
    w = torch.empty((B, R * k), dtype=torch.float32, device=adj.device)
    with torch.cuda.device(adj.device):
        _lib.call("igcn_gdc_topk_emit", _lib.ptr(ppr), B, R, k, _lib.ptr(src), _lib.ptr(dst), _lib.ptr(w), _lib.stream(), tag="gdc_topk_emit",
                  nbytes=8 * B * R * R + 12 * B * R * k)
    return src, dst, w


def knn_connectivity_device(z: torch.Tensor, knn: int = 5):
    """|corr| of the rows of z (B,R,T), top-`knn` per row without the diagonal, symmetrised by max (synthetic.knn_connectivity)."""
    zc = z - z.mean(-1, keepdim=True)
    zn = zc / zc.norm(dim=-1, keepdim=True).clamp_min(1e-30)
    c = (zn @ zn.transpose(1, 2)).abs()
    R = c.shape[1]
    c = c * (1.0 - torch.eye(R, dtype=c.dtype, device=c.device))
    idx = c.topk(knn, dim=-1).indices
    m = torch.zeros_like(c).scatter_(-1, idx, c.gather(-1, idx))
    return torch.maximum(m, m.transpose(1, 2))


class DeviceSubjectSet(object):
    """A dataset of equally sized brain graphs held as DEVICE tensors (the SubjectSet field list)."""

    def __init__(self, **fields):
        for k_, v in fields.items():
            setattr(self, k_, v)
        self.n, self.rois = int(self.x.shape[0]), int(self.x.shape[1])
        self.eg = int(self.edge_src.shape[1])
        self.device = self.x.device

    def __len__(self):
        return self.n

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.__dict__.values() if torch.is_tensor(t))

    @staticmethod
    def generate(num, rois=90, n_snps=54, seed=0, device="cuda", first_id=0, num_classes=3, num_regr=3, feat=3, top_k=3, chunk=2048,
                 series=16) -> "DeviceSubjectSet":
        """`num` synthetic subjects generated and GDC-preprocessed on `device`, `chunk` subjects at a time (the fp64 PPR inverse of a
        chunk is chunk * R^2 * 8 bytes)."""
        dev = torch.device(device)
        g = torch.Generator(device=dev)
        g.manual_seed(int(seed) * 1000003 + int(first_id))
        xs, srcs, dsts, ws, snps, ys, cls, clu = [], [], [], [], [], [], [], []
        for lo in range(0, num, chunk):
            n = min(chunk, num - lo)
            x = torch.rand((n, rois, feat), generator=g, device=dev)
            z = torch.randn((n, rois, series), generator=g, device=dev, dtype=torch.float64)
            s, d, w = gdc_topk_device(knn_connectivity_device(z), k=top_k)
            xs.append(x); srcs.append(s); dsts.append(d); ws.append(w)
            snps.append(torch.randint(0, 3, (n, n_snps), generator=g, device=dev).float() * 0.5)
            ys.append(torch.randint(0, num_classes, (n,), generator=g, device=dev))
            cls.append(torch.rand((n, num_regr), generator=g, device=dev))
            clu.append(torch.randint(0, 2, (n,), generator=g, device=dev))
        x = torch.cat(xs)
        return DeviceSubjectSet(x=x, edge_src=torch.cat(srcs), edge_dst=torch.cat(dsts), edge_attr=torch.cat(ws), snps_feat=torch.cat(snps),
                                y=torch.cat(ys), clini_score=torch.cat(cls), tsne_fdim=x[:, :, feat - 1].contiguous(), clust_y=torch.cat(clu),
                                sbjID=torch.arange(first_id, first_id + num, device=dev))


def collate_device(ds: DeviceSubjectSet, idx: torch.Tensor, out: Batch = None) -> Batch:
    """Minibatch `idx` (int64 device tensor) of a DeviceSubjectSet: device gathers + igcn_collate_csr; no host transfer.
    `out`: a Batch of the same size whose buffers are overwritten in place (the static inputs of a captured training step)."""
    dev = ds.device
    B, R, eg = int(idx.numel()), ds.rois, ds.eg
    E = B * eg
    pick = lambda t: t.index_select(0, idx)
    d = dict(x=pick(ds.x), snps_feat=pick(ds.snps_feat), y=pick(ds.y), clini_score=pick(ds.clini_score), tsne_fdim=pick(ds.tsne_fdim),
             clust_y=pick(ds.clust_y), sbjID=pick(ds.sbjID), edge_src=pick(ds.edge_src).view(-1), edge_dst=pick(ds.edge_dst).view(-1),
             edge_attr=pick(ds.edge_attr).view(-1))
    if out is not None:
        if out._num_graphs != B or out._csr.E != E:
            raise RuntimeError("collate_device(out=...): static batch has B=%d,E=%d, new batch B=%d,E=%d" % (out._num_graphs, out._csr.E, B, E))
        for k_, v in d.items():
            out._raw[k_].copy_(v)
        return Batch._finish(out._raw, out._gptr, B, R, E, eg, dev, out)
    gptr = torch.arange(0, (B + 1) * eg, eg, dtype=torch.int64, device=dev)
    return Batch._finish(d, gptr, B, R, E, eg, dev, None)
