"""Seeded synthetic ADNI-shaped subjects and GO hierarchies.

The real ADNI imaging/SNP data and the PANTHER/CTD GO files are not available
offline, so the hot path is exercised on synthetic inputs whose SHAPE and
SPARSITY PATTERN restate what the reference's host pipeline produces:

  * brain graph: symmetric kNN(5) |corr| connectivity (stands in for
    data/brain_image/knn/5/corr_data.mat, main.py:194) -> graph diffusion as in
    util_gdc.py:7-14 (PPR, alpha=0.05) -> keep top-k=3 per COLUMN and
    column-normalise (util_gdc.py:25-31, main.py:45) -> dense->COO in row-major
    order (util_gdc.py:84-86).  Result: exactly k in-edges per node, weights in
    (0,1] summing to 1 per target, a self loop on (almost) every node.
  * per-subject fields as built at sgcn_data.py:257-288.
  * GO DAG with the level-sorted (deepest first) node order and `pool_dim`
    contract of snps_graph.py:273-289, GO x SNP incidence with the all-ones root
    row (snps_graph.py:247-248).

Everything here is host-side numpy; it is input generation, not the hot path.
"""
from __future__ import annotations

import numpy as np

ADNI_POOL = [20, 15, 10, 8, 1]          # ADNI-shaped GO, G=54
LARGE_POOL = [1200, 500, 200, 99, 1]    # config 3: G=2000


def gdc_topk(adj: np.ndarray, alpha: float = 0.05, k: int = 3) -> np.ndarray:
    """Batched restatement of util_gdc.py:7-14 + 25-31. adj: (B,R,R) symmetric, positive row sums."""
    B, R, _ = adj.shape
    dinv = 1.0 / np.sqrt(adj.sum(-1))
    H = dinv[:, :, None] * adj * dinv[:, None, :]
    ppr = alpha * np.linalg.inv(np.eye(R)[None] - (1.0 - alpha) * H)
    order = np.argsort(ppr, axis=1)                   # per column, ascending over rows
    drop = order[:, : R - k, :]                       # the R-k smallest rows of every column
    b_idx = np.arange(B)[:, None, None]
    c_idx = np.arange(R)[None, None, :]
    ppr[b_idx, drop, c_idx] = 0.0
    norm = ppr.sum(1)
    norm[norm <= 0] = 1.0
    return ppr / norm[:, None, :]


def knn_connectivity(rng_list, R: int, knn: int = 5, t: int = 16) -> np.ndarray:
    """|corr| of R random length-t series, top-`knn` per row (no diagonal), symmetrised by max."""
    B = len(rng_list)
    out = np.zeros((B, R, R))
    for b, rng in enumerate(rng_list):
        z = rng.standard_normal((R, t))
        c = np.abs(np.corrcoef(z))
        np.fill_diagonal(c, 0.0)
        keep = np.argsort(-c, axis=1)[:, :knn]
        m = np.zeros_like(c)
        rows = np.arange(R)[:, None]
        m[rows, keep] = c[rows, keep]
        out[b] = np.maximum(m, m.T)
    return out


def make_subjects(num: int, rois: int = 90, n_snps: int = 54, seed: int = 0, first_id: int = 0,
                  num_classes: int = 3, num_regr: int = 3, feat: int = 3, top_k: int = 3):
    """Returns a dict of packed per-subject host arrays (the layout `Batch` collates from):

      x (num,R,F0) f32 | edge_ptr (num+1,) i64 | edge_src/edge_dst (E,) i64 LOCAL ids, row-major COO |
      edge_attr (E,) f32 | snps_feat (num,S) f32 | y (num,) i64 | clini_score (num,num_regr) f32 |
      tsne_fdim (num,R) f32 | clust_y (num,) i64 | sbjID (num,) i64
    """
    rngs = [np.random.default_rng([seed, first_id + i]) for i in range(num)]
    x = np.stack([r.random((rois, feat)) for r in rngs]).astype(np.float32)
    conn = knn_connectivity(rngs, rois)
    diff = gdc_topk(conn, k=top_k)
    snps = np.stack([r.integers(0, 3, size=n_snps) for r in rngs]).astype(np.float32) * 0.5
    y = np.array([r.integers(0, num_classes) for r in rngs], dtype=np.int64)
    clini = np.stack([r.random(num_regr) for r in rngs]).astype(np.float32)
    clust = np.array([r.integers(0, 2) for r in rngs], dtype=np.int64)
    src, dst, w, ptr = [], [], [], [0]
    for b in range(num):
        a32 = diff[b].astype(np.float32)
        s, d = np.nonzero(a32)                         # row-major == scipy coo_matrix(dense) order
        src.append(s.astype(np.int64))
        dst.append(d.astype(np.int64))
        w.append(a32[s, d])
        ptr.append(ptr[-1] + s.size)
    return dict(
        x=x, edge_ptr=np.asarray(ptr, dtype=np.int64),
        edge_src=np.concatenate(src), edge_dst=np.concatenate(dst),
        edge_attr=np.concatenate(w).astype(np.float32),
        snps_feat=snps, y=y, clini_score=clini, tsne_fdim=np.ascontiguousarray(x[:, :, feat - 1]),
        clust_y=clust, sbjID=np.arange(first_id, first_id + num, dtype=np.int64),
        rois=rois, n_snps=n_snps,
    )


def make_go_hierarchy(pool=None, n_snps: int = 54, seed: int = 0):
    """GO DAG in the reference layout. Returns (adj (G,G) 0/1 with adj[child,parent]=1,
    go_snps (G,S) 0/1, pool_dim [[p0..p4]]).  The model receives A = adj.T
    (kernel/train_eval_sgcn_img_snps.py:69) so A[parent, child] = 1."""
    pool = list(ADNI_POOL if pool is None else pool)
    rng = np.random.default_rng([seed, 777])
    G = int(sum(pool))
    start = np.concatenate([[0], np.cumsum(pool)])
    adj = np.zeros((G, G), dtype=np.float32)
    for lvl in range(len(pool) - 1):
        lo, hi = start[lvl + 1], start[lvl + 2]          # the next (shallower) level
        for c in range(start[lvl], start[lvl + 1]):
            npar = min(int(rng.integers(1, 4)), hi - lo)
            parents = rng.choice(np.arange(lo, hi), size=npar, replace=False)
            adj[c, parents] = 1.0
    go_snps = (rng.random((G, n_snps)) < (3.0 / G)).astype(np.float32)
    go_snps[G - 1, :] = 1.0                              # root row (snps_graph.py:247-248)
    return adj, go_snps, [pool]
