#!/usr/bin/env python
"""cuobjdump -sass of libpvsg_sm100.so -> per-kernel counts of the mnemonics that prove (or disprove) a Blackwell-native
kernel (B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), HMMA / LDSM (legacy).
  python tools/sass_mnemonics.py [out.json]"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(['cuobjdump', '-sass', os.path.join(ROOT, 'openpvsg_b200', 'libpvsg_sm100.so')], capture_output=True, text=True).stdout
WATCH = ('UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'HMMA', 'LDSM', 'LDGSTS', 'SYNCS')
out, cur = {}, None
for line in txt.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(anonymous namespace\)::', '', name)
        cur = out.setdefault(re.sub(r'\(.*', '', name), collections.Counter())
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur is not None:
        op = m.group(1)
        cur['instructions'] += 1
        for w in WATCH:
            if op.startswith(w):
                cur[w] += 1
res = {k: dict(v) for k, v in sorted(out.items()) if any(w in v for w in WATCH[:9])}
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02_sass_mnemonics.json')
json.dump(dict(source='cuobjdump -sass openpvsg_b200/libpvsg_sm100.so (tools/sass_mnemonics.py); kernels containing tensor-core / TMA instructions',
               kernels=res), open(dst, 'w'), indent=1)
for k, v in res.items():
    print(k, {w: v[w] for w in WATCH if w in v})
