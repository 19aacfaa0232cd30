"""Device time of one frame at batch 1 (graph replay) and where it goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import openpvsg_b200 as pv
from openpvsg_b200 import configs, engine, synthetic as syn
det = pv.build_detector(configs.mask2former_r50(True))
det.load_state_dict(syn.mask2former_state_dict(seed=0))
det.cuda()
engine.enable_cuda_graph(det)
H, W = 720, 1280
meta = syn.frame_meta(H, W)
f = syn.synthetic_frame(0, H, W).pin_memory()
r = engine.get_runner(det, meta, True, batch=1)
for _ in range(3):
    r.run(f)
g = r.lane_graph[0]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.replay()
e1.record(); torch.cuda.synchronize()
print('graph replay batch 1: %.3f ms' % (e0.elapsed_time(e1) / 20))
import time
t0 = time.perf_counter()
for _ in range(20):
    r.run(f)
print('run() end to end: %.3f ms' % ((time.perf_counter() - t0) / 20 * 1e3))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=60))
