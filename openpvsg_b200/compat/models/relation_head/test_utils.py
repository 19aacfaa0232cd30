from openpvsg_b200.relation_head import generate_pairwise_results, generate_results, pick_top_pairs_eval  # noqa: F401
