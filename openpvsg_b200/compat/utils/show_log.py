"""utils/show_log.py::save_metrics_to_csv -- one CSV row per evaluated model (tools/rel_test.py:112-113)."""
import csv
import os

import numpy as np


def save_metrics_to_csv(final_metrics, pair_recall_list, K_values, csv_file_path, model_name):
    new_file = not os.path.isfile(csv_file_path)
    header = ['Model', 'Pair Recall'] + [f'R/mR@{K}' for K in K_values] + [f'wR/wmR@{K}' for K in K_values]
    pct = lambda v: f'{100 * v:.2f}'  # noqa: E731
    row = [model_name, pct(np.array(pair_recall_list).mean())]
    row += [f"{pct(final_metrics[K]['recall'])}/{pct(final_metrics[K]['mean_recall'])}" for K in K_values]
    row += [f"{pct(final_metrics[K]['weak_recall'])}/{pct(final_metrics[K]['weak_mean_recall'])}" for K in K_values]
    with open(csv_file_path, mode='a', newline='') as f:
        writer = csv.writer(f)
        if new_file:
            writer.writerow(header)
        writer.writerow(row)
