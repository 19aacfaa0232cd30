from openpvsg_b200.relation_head import PositionalEncoding, TemporalTransformer  # noqa: F401
