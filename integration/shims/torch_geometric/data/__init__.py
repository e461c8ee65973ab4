"""torch_geometric.data names the path imports (batch.py:3, dataloader.py:4, kernel/train_eval_sgcn_img_snps.py:11)."""
from igcn_b200.data import Batch, Data, DataLoader  # noqa: F401

DenseDataLoader = DataLoader        # imported (never used) by kernel/train_eval_sgcn_img_snps.py:11
InMemoryDataset = object            # base class of the reference's dataset wrappers (outside the hot path)
Dataset = object
