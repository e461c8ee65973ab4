"""Drop-in batching layer: `Data`, `Batch`, `DataLoader` (reference: batch.py, dataloader.py; PyG `Data`).

The reference collates on the CPU in Python (B x ~11 keys of tensor adds and appends, then 11 torch.cat,
batch.py:51-110) and ships the result -- including a dense per-subject `A` the model never reads -- to the
device.  Here a dataset is held once as packed, pinned host arrays (`SubjectSet`); a mini-batch is a gather
of per-subject slabs into one pinned staging buffer, one H2D copy per field, and ONE CUDA kernel
(`igcn_collate_csr`) that emits, on the device:

  * `edge_index` (2,E) i64 and `batch` (N,) i64 -- bit exact with Batch.from_data_list, and
  * the target-sorted CSR + source-sorted transposed index the fused SGCN kernels consume.

`Batch` exposes the attributes the reference training loop reads (kernel/train_eval_sgcn_img_snps.py:511-548):
x, edge_index, edge_attr, batch, y, snps_feat, clini_score, tsne_fdim, clust_y, sbjID, num_graphs, .to().
"""
from __future__ import annotations

import weakref
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib


class Data(object):
    """Minimal stand-in for torch_geometric.data.Data (attribute bag of tensors)."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, **kwargs).items():
            if v is not None:
                setattr(self, k, v)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if not k.startswith("_") and v is not None]

    def __getitem__(self, k):
        return getattr(self, k, None)

    def __setitem__(self, k, v):
        setattr(self, k, v)

    def __contains__(self, k):
        return k in self.keys

    @property
    def num_nodes(self):
        x = getattr(self, "x", None)
        return None if x is None else x.size(0)

    def to(self, device, non_blocking=False):
        for k in self.keys:
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self


class GraphCSR(object):
    """Device-resident sparse structure of one collated batch (all i32 / f32)."""
    __slots__ = ("rowptr_t", "csr_src", "csr_perm", "csr_w", "rowptr_s", "csc_pos", "max_eg", "B", "R", "E")

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


# ---- batch-structure registry: lets the PyG-signature operators (pyg.py), which receive bare `edge_index` / `batch` tensors, find
#      the structure of a collation this package produced (graphs of R nodes, CSR already built) without a host round trip -------------
_structures = {}


def _structure_key(t):
    return (t.data_ptr(), tuple(t.shape), str(t.device))


def register_structure(csr, *tensors):
    for t in tensors:
        if t is not None:
            _structures[_structure_key(t)] = (weakref.ref(t), t._version, csr)
    while len(_structures) > 64:
        _structures.pop(next(iter(_structures)))


def lookup_structure(t):
    """GraphCSR registered for this `edge_index` / `batch` tensor by Batch (same storage and shape), or None.  An entry is valid
    only while the tensor it was registered for is alive and unedited: once that tensor is freed the allocator may hand its
    address to an unrelated tensor of the same shape (the next batch of a fixed-size loader), which must not inherit the structure."""
    if t is None:
        return None
    key = _structure_key(t)
    e = _structures.get(key)
    if e is None:
        return None
    owner = e[0]()
    if owner is None or owner._version != e[1]:
        _structures.pop(key, None)
        return None
    return e[2]


class SubjectSet(object):
    """A dataset of equally-sized brain graphs as packed host arrays (see synthetic.make_subjects for the
    field list).  Arrays are converted once to torch tensors in pinned memory when CUDA is available."""

    FIELDS = ("x", "snps_feat", "y", "clini_score", "tsne_fdim", "clust_y", "sbjID")

    def __init__(self, packed: dict, pin: Optional[bool] = None):
        pin = torch.cuda.is_available() if pin is None else pin
        self.rois = int(packed["x"].shape[1])
        self.n = int(packed["x"].shape[0])
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
        self.x = t(packed["x"], torch.float32)
        self.snps_feat = t(packed["snps_feat"], torch.float32)
        self.y = t(packed["y"], torch.int64)
        self.clini_score = t(packed["clini_score"], torch.float32)
        self.tsne_fdim = t(packed["tsne_fdim"], torch.float32)
        self.clust_y = t(packed["clust_y"], torch.int64)
        self.sbjID = t(packed["sbjID"], torch.int64)
        self.edge_ptr = t(packed["edge_ptr"], torch.int64)
        self.edge_src = t(packed["edge_src"], torch.int32)      # LOCAL ids; i32 halves the H2D bytes
        self.edge_dst = t(packed["edge_dst"], torch.int32)
        self.edge_attr = t(packed["edge_attr"], torch.float32)
        # host contract of igcn_collate_csr (validated once here; the kernel indexes shared memory with these ids)
        if self.x.dim() != 3:
            raise ValueError("SubjectSet: x must be (subjects, rois, features), got %s" % (tuple(self.x.shape),))
        ne = int(self.edge_ptr[-1]) if self.edge_ptr.numel() else 0
        if self.edge_ptr.numel() != self.n + 1 or int(self.edge_ptr[0]) != 0 or bool((self.edge_ptr[1:] < self.edge_ptr[:-1]).any()) \
                or not (self.edge_src.numel() == self.edge_dst.numel() == self.edge_attr.numel() == ne):
            raise ValueError("SubjectSet: edge_ptr must be a non-decreasing prefix over %d subjects covering all %d edges"
                             % (self.n, self.edge_src.numel()))
        if ne and (int(self.edge_src.min()) < 0 or int(self.edge_src.max()) >= self.rois or
                   int(self.edge_dst.min()) < 0 or int(self.edge_dst.max()) >= self.rois):
            raise ValueError("SubjectSet: edge endpoints must be LOCAL node ids in [0, %d)" % self.rois)
        counts = self.edge_ptr[1:] - self.edge_ptr[:-1]
        self.edge_counts = counts
        self.max_eg = int(counts.max()) if self.n else 0
        self.uniform_eg = int(counts[0]) if self.n and bool((counts == counts[0]).all()) else None
        if pin:
            for k in self.FIELDS + ("edge_src", "edge_dst", "edge_attr"):
                setattr(self, k, getattr(self, k).pin_memory())

    def __len__(self):
        return self.n

    @staticmethod
    def from_data_list(data_list: Sequence[Data], pin: Optional[bool] = None) -> "SubjectSet":
        """Pack a list of reference-style `Data` objects (sgcn_data.py:282-288 field names)."""
        ptr, src, dst, w = [0], [], [], []
        for d in data_list:
            ei = d.edge_index.cpu().numpy()
            src.append(ei[0])
            dst.append(ei[1])
            w.append(d.edge_attr.cpu().numpy())
            ptr.append(ptr[-1] + ei.shape[1])
        get = lambda name: [getattr(d, name).cpu().numpy() for d in data_list]
        n = len(data_list)
        cat1 = lambda xs: np.concatenate([np.asarray(v).reshape(-1) for v in xs]) if n else np.zeros((0,))
        packed = dict(
            x=np.stack(get("x")), edge_ptr=np.asarray(ptr, np.int64), edge_src=cat1(src), edge_dst=cat1(dst),
            edge_attr=cat1(w), snps_feat=np.stack([v.reshape(-1) for v in get("snps_feat")]),
            y=cat1(get("y")), clini_score=np.stack([v.reshape(-1) for v in get("clini_score")]),
            tsne_fdim=np.stack([v.reshape(-1) for v in get("tsne_fdim")]), clust_y=cat1(get("clust_y")),
            sbjID=cat1(get("sbjID")))
        return SubjectSet(packed, pin)


class Batch(Data):
    """One big disconnected graph on the device + its CSR. Mirrors the reference Batch surface."""

    def __init__(self):
        super().__init__()
        self._num_graphs = 0
        self._csr = None
        self.rois = 0

    @property
    def num_graphs(self):
        # the reference does batch[-1].item()+1 (batch.py:188-191, a device sync); the count is known here
        return self._num_graphs

    @property
    def csr(self) -> GraphCSR:
        return self._csr

    def to(self, device, non_blocking=False):
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())       # 'cuda' means the current device, as in torch
        if self.x is not None and self.x.device == dev:
            return self
        raise RuntimeError("igcn_b200.Batch is collated on its CUDA device; it cannot be moved to %s" % dev)

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def collate(ss: SubjectSet, idx, device, staging: Optional[dict] = None, out: Optional["Batch"] = None) -> "Batch":
        """Gather subjects `idx` (host, pinned), copy to `device`, collate + build CSR with one kernel.
        `out`: a Batch of the same (B, E) whose device buffers are overwritten in place (static buffers of a
        captured CUDA graph)."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("igcn_b200.Batch.collate needs a CUDA device (no CPU fallback)")
        idx_t = torch.as_tensor(idx, dtype=torch.int64)
        B, R = int(idx_t.numel()), ss.rois
        h = {}
        pin = ss.x.is_pinned()
        if staging is not None and staging.get("_event") is not None:
            staging["_event"].synchronize()      # the previous batch's H2D copies have left the staging buffers

        def take(name, src):
            out = None
            if staging is not None:
                buf = staging.get(name)
                if buf is not None and buf.shape[0] >= B and buf.shape[1:] == src.shape[1:]:
                    out = buf[:B]
            if out is None:
                out = torch.empty((B,) + tuple(src.shape[1:]), dtype=src.dtype, pin_memory=pin)
                if staging is not None:
                    staging[name] = out
            torch.index_select(src, 0, idx_t, out=out)
            return out

        for name in SubjectSet.FIELDS:
            h[name] = take(name, getattr(ss, name))
        counts = ss.edge_counts[idx_t]
        gptr = torch.zeros(B + 1, dtype=torch.int64, pin_memory=pin)
        torch.cumsum(counts, 0, out=gptr[1:])
        E = int(gptr[-1])
        if ss.uniform_eg is not None:
            eg = ss.uniform_eg
            for name in ("edge_src", "edge_dst", "edge_attr"):
                h[name] = take(name, getattr(ss, name).view(ss.n, eg)).view(-1)
        else:
            ep = ss.edge_ptr
            sel = torch.cat([torch.arange(int(ep[g]), int(ep[g + 1])) for g in idx_t.tolist()]) if B else torch.zeros(0, dtype=torch.int64)
            for name in ("edge_src", "edge_dst", "edge_attr"):
                src = getattr(ss, name)
                buf = torch.empty(E, dtype=src.dtype, pin_memory=pin)
                torch.index_select(src, 0, sel, out=buf)
                h[name] = buf
        max_eg = int(counts.max()) if B else 0
        if out is not None:
            if out._num_graphs != B or out._csr.E != E or out._csr.max_eg < max_eg:
                raise RuntimeError("Batch.collate(out=...): static batch has B=%d,E=%d, new batch B=%d,E=%d"
                                   % (out._num_graphs, out._csr.E, B, E))
            d = out._raw
            for k, v in h.items():
                d[k].copy_(v, non_blocking=True)
            out._gptr.copy_(gptr, non_blocking=True)
            d_gptr = out._gptr
        else:
            d = {k: v.to(dev, non_blocking=True) for k, v in h.items()}
            d_gptr = gptr.to(dev, non_blocking=True)
        if staging is not None:
            ev = staging.get("_event") or torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            staging["_event"] = ev
        return Batch._finish(d, d_gptr, B, R, E, max_eg, dev, out)

    @staticmethod
    def _finish(d, d_gptr, B, R, E, max_eg, dev, out=None) -> "Batch":
        N = B * R
        i32 = dict(dtype=torch.int32, device=dev)
        if out is not None:
            b, csr = out, out._csr
        else:
            b = Batch()
            b.edge_index = torch.empty((2, E), dtype=torch.int64, device=dev)
            b.batch = torch.empty(N, dtype=torch.int64, device=dev)
            csr = GraphCSR(rowptr_t=torch.empty(N + 1, **i32), csr_src=torch.empty(E, **i32), csr_perm=torch.empty(E, **i32),
                           csr_w=torch.empty(E, dtype=torch.float32, device=dev), rowptr_s=torch.empty(N + 1, **i32),
                           csc_pos=torch.empty(E, **i32), max_eg=max_eg, B=B, R=R, E=E)
            if B == 0:
                csr.rowptr_t.zero_()
                csr.rowptr_s.zero_()
        L = _lib.lib()
        with torch.cuda.device(dev):
            _lib.call("igcn_collate_csr", _lib.ptr(d_gptr), _lib.ptr(d["edge_src"]), _lib.ptr(d["edge_dst"]), _lib.ptr(d["edge_attr"]),
                                    B, R, E, max_eg, _lib.ptr(b.edge_index), _lib.ptr(b.batch), _lib.ptr(csr.rowptr_t),
                                    _lib.ptr(csr.csr_src), _lib.ptr(csr.csr_perm), _lib.ptr(csr.csr_w), _lib.ptr(csr.rowptr_s),
                                    _lib.ptr(csr.csc_pos), _lib.stream())
        if out is not None:
            return b                      # every attribute already aliases the static buffers
        b._raw, b._gptr = d, d_gptr
        b.x = d["x"].view(N, d["x"].shape[-1])
        b.edge_attr = d["edge_attr"]
        b.snps_feat = d["snps_feat"]
        b.y = d["y"]
        b.clini_score = d["clini_score"].reshape(-1)
        b.tsne_fdim = d["tsne_fdim"]
        b.clust_y = d["clust_y"]
        b.sbjID = d["sbjID"]
        b._num_graphs, b._csr, b.rois = B, csr, R
        register_structure(csr, b.edge_index, b.batch)
        return b

    @staticmethod
    def from_data_list(data_list: List[Data], follow_batch=(), device=None) -> "Batch":
        """Reference signature (batch.py:24).  Packs the list on the host, collates on `device`."""
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        ss = SubjectSet.from_data_list(data_list)
        return Batch.collate(ss, np.arange(len(ss)), dev)

    @staticmethod
    def from_device_tensors(x, edge_index, edge_attr, rois, **extra) -> "Batch":
        """Wrap an already collated (reference-style) batch that lives on the device: builds the CSR from the
        global int64 edge_index.  Needs one host read of the maximum per-graph edge count."""
        _lib.require_cuda(x, edge_index, edge_attr)
        dev = x.device
        N, E = x.shape[0], edge_index.shape[1]
        B = N // rois
        i32 = dict(dtype=torch.int32, device=dev)
        ei = edge_index.contiguous()
        if N != B * rois:
            raise ValueError("from_device_tensors: %d nodes is not a multiple of rois=%d" % (N, rois))
        gid = torch.div(ei[0], rois, rounding_mode="floor")
        if E:
            # one host read validates the kernel's contract (edges grouped by graph, no cross-graph edge, ids in range) together
            # with the maximum per-graph edge count the launch needs anyway
            gd = torch.div(ei[1], rois, rounding_mode="floor")
            bad = (gid != gd).any() | (gid[1:] < gid[:-1]).any() | (ei.min() < 0) | (ei.max() >= N)
            stats = torch.stack([torch.bincount(gid.clamp(0, max(B - 1, 0)), minlength=B).max(), bad.to(torch.int64)]).tolist()
            if stats[1]:
                raise ValueError("from_device_tensors: edge_index must be a collated batch of %d-node graphs (edges grouped by graph, "
                                 "no edge between graphs, ids in [0, %d))" % (rois, N))
            max_eg = int(stats[0])
        else:
            max_eg = 0
        eptr = torch.empty(B + 1, **i32)
        csr = GraphCSR(rowptr_t=torch.empty(N + 1, **i32), csr_src=torch.empty(E, **i32), csr_perm=torch.empty(E, **i32),
                       csr_w=torch.empty(E, dtype=torch.float32, device=dev), rowptr_s=torch.empty(N + 1, **i32),
                       csc_pos=torch.empty(E, **i32), max_eg=max_eg, B=B, R=rois, E=E)
        L = _lib.lib()
        with torch.cuda.device(dev):
            _lib.call("igcn_csr_from_edge_index", _lib.ptr(ei), _lib.ptr(edge_attr.contiguous()), B, rois, E, max_eg, _lib.ptr(eptr),
                                            _lib.ptr(csr.rowptr_t), _lib.ptr(csr.csr_src), _lib.ptr(csr.csr_perm),
                                            _lib.ptr(csr.csr_w), _lib.ptr(csr.rowptr_s), _lib.ptr(csr.csc_pos), _lib.stream())
        b = Batch()
        b.x, b.edge_index, b.edge_attr = x, ei, edge_attr
        b.batch = torch.arange(B, device=dev).repeat_interleave(rois)
        for k, v in extra.items():
            setattr(b, k, v)
        b._num_graphs, b._csr, b.rois = B, csr, rois
        register_structure(csr, b.edge_index, b.batch)
        return b


class DataLoader(object):
    """Reference surface: DataLoader(dataset, batch_size, shuffle) (dataloader.py:11-48); iterating yields
    device-resident `Batch` objects.  `dataset` is a SubjectSet or a list of `Data`."""

    def __init__(self, dataset, batch_size=1, shuffle=False, follow_batch=(), device=None, generator=None, drop_last=False, **kwargs):
        self.dataset = dataset if isinstance(dataset, SubjectSet) else SubjectSet.from_data_list(list(dataset))
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.generator = generator
        self._staging = {}

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.generator) if self.shuffle else torch.arange(n)
        for s in range(0, n, self.batch_size):
            idx = order[s:s + self.batch_size]
            if self.drop_last and idx.numel() < self.batch_size:
                break
            yield Batch.collate(self.dataset, idx, self.device, self._staging)
