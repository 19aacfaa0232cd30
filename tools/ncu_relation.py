#!/usr/bin/env python
"""One relation-head forward (200 tubes x 128 frames) between cudaProfilerStart/Stop."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openpvsg_b200 import relation_head as rh, synthetic as syn  # noqa: E402

dev = torch.device('cuda:0')
sds = syn.relation_state_dicts(seed=1)
feats = torch.randn(200, 128, 256, generator=torch.Generator().manual_seed(0)).to(dev)
mods = [rh.ObjectEncoder(256), rh.ObjectEncoder(256), rh.PairProposalNetwork(256, 1024), rh.TemporalTransformer(512, 57)]
for m, k in zip(mods, ('subject_encoder', 'object_encoder', 'pair_proposal_model', 'relation_model')):
    m.load_state_dict(sds[k])
    m.to(dev)
for _ in range(2):
    rh.relation_forward(*mods, feats, 100)
torch.cuda.synchronize()
torch.cuda.profiler.start()
rh.relation_forward(*mods, feats, 100)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
