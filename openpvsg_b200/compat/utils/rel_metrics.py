from openpvsg_b200.rel_eval import calculate_final_metrics, calculate_iou, calculate_pair_recall_at_k  # noqa: F401
