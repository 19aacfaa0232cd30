from openpvsg_b200.relation_head import HandcraftedFilter, Learnable1DConv  # noqa: F401
