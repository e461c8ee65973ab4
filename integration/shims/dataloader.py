"""Shadow of the reference's dataloader.py (dataloader.py:11-48): put `integration/shims` BEFORE the reference directory on
sys.path and `from dataloader import DataLoader` in kernel/train_eval_sgcn_img_snps.py:21 resolves here."""
from igcn_b200.data import Batch, Data, DataLoader  # noqa: F401
