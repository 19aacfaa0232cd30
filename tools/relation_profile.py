"""Warm per-kernel GPU times of one relation-head forward (200 tubes x 128 frames, top-100 pairs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from openpvsg_b200 import relation_head as rh, synthetic as syn
sds = syn.relation_state_dicts(seed=1)
mods = [rh.ObjectEncoder(256), rh.ObjectEncoder(256), rh.PairProposalNetwork(256, 1024), rh.TemporalTransformer(512, 57)]
for m, k in zip(mods, ('subject_encoder', 'object_encoder', 'pair_proposal_model', 'relation_model')):
    m.load_state_dict(sds[k]); m.cuda()
fd = torch.randn(200, 128, 256, generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(5):
    rh.relation_forward(*mods, fd, 100)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        rh.relation_forward(*mods, fd, 100)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=20, max_name_column_width=70))
