"""Shadow of kernel/sgcn_img_snp.py: `from kernel.sgcn_img_snp import SGCN_GCN_IMGSNP` (train_eval_sgcn_img_snps.py:32)."""
from igcn_b200.img_snp_model import SGCN_GCN_IMGSNP  # noqa: F401
