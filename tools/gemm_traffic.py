#!/usr/bin/env python
"""Summarise an ncu pass (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum over the
gemm_tc_kernel launches of one frame-graph replay) into profiles/<tag>_gemm_traffic.json."""
import collections
import csv
import json
import sys

path, out = sys.argv[1], sys.argv[2]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 20
lines = [l for l in open(path) if not l.startswith('==')]
per = collections.defaultdict(dict)
unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3}
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', '')) * unit.get(row['Metric Unit'], 1.0)
    per[row['ID']][row['Metric Name']] = v
n = len(per)
rd = sum(d.get('dram__bytes_read.sum', 0.0) for d in per.values())
wr = sum(d.get('dram__bytes_write.sum', 0.0) for d in per.values())
us = sum(d.get('gpu__time_duration.sum', 0.0) for d in per.values())
res = dict(source='ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_tc_kernel '
                  f'--profile-from-start off python tools/ncu_frame.py {frames} (one replay of the {frames}-frame 720p graph)',
           launches=n, dram_read_bytes=rd, dram_write_bytes=wr, total_us_under_ncu=round(us, 1),
           avg_bytes_per_launch=round((rd + wr) / max(n, 1)), frames=frames,
           dram_bytes_per_frame=round((rd + wr) / frames))
json.dump(res, open(out, 'w'), indent=1)
print(res)
