from openpvsg_b200.relation_head import ObjectEncoder, PairProposalNetwork, VanillaModel  # noqa: F401
