#!/usr/bin/env python
"""Condense ncu outputs into small JSON files for profiles/.

  ncu_summary.py launches <launches.csv> <out.json> [note]     # per-kernel time shares
  ncu_summary.py full <report.ncu-rep> <out.json>              # key metrics of a --set full capture
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio']


def short(name):
    s = re.sub(r'^void\s+', '', name.strip())
    s = s.replace('<unnamed>::', '').replace('(anonymous namespace)::', '')
    m = re.match(r'([A-Za-z_][A-Za-z0-9_:]*)', s)
    base = m.group(1) if m else s[:40]
    t = re.match(r'[A-Za-z_][A-Za-z0-9_:]*<(.*?)>\(', s)
    targs = re.sub(r'\((int|bool)\)', '', t.group(1)) if t else ''
    return base + (f'<{targs}>' if targs and len(targs) < 40 else '')


def launches(path, out, note=''):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(row['Metric Unit'], 1.0)
        k = short(row['Kernel Name'])
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    ks = [dict(kernel=k, launches=n, total_us=round(t, 1), share=round(t / tot, 4))
          for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])]
    json.dump(dict(source=note, total_us=round(tot, 1), launches=sum(k['launches'] for k in ks), kernels=ks),
              open(out, 'w'), indent=1)
    for k in ks[:12]:
        print(k)


def full(rep, out):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = dict(kernel=short(row[hdr.index('Kernel Name')]))
        for k in KEYS:
            if k in hdr:
                d[k] = f'{row[hdr.index(k)]} {units[hdr.index(k)]}'.strip()
        res.append(d)
    json.dump(res, open(out, 'w'), indent=1)
    for d in res:
        print(d)


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '')
    else:
        full(sys.argv[2], sys.argv[3])
