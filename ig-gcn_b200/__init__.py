"""igcn_b200 -- B200-native (sm_100a) implementation of the IG-GCN graph-convolution hot path.

Public surface mirrors the reference (Houliang-Zhou/IG-GCN): `Data`/`Batch`/`DataLoader` (batch.py,
dataloader.py), PyG-style operators, `SGCN_GCN_IMGSNP`, `Gene_ontology_network`.  All compute goes
through hand-written CUDA kernels behind the C ABI in include/igcn_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"
