"""`datasets` of the reference's tools: the relation dataset (datasets/datasets/pvsg_relation.py) and the dataset
builder; disk-backed VPS datasets stay with the reference (see ../README.md)."""
from openpvsg_b200.relation_set import PVSGRelationDataset  # noqa: F401
from .datasets import builder  # noqa: F401
