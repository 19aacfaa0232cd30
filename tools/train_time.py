"""Time one training step (forward_train + backward + AdamW) of the VPS detector at the reference's training shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import openpvsg_b200 as pv
from openpvsg_b200 import configs, synthetic as syn

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, W, T = 384, 480, 2
det = pv.build_detector(configs.mask2former_r50(True))
det.load_state_dict(syn.mask2former_state_dict(seed=3))
det.cuda()
det.panoptic_head.train_cfg = dict(num_points=12544, oversample_ratio=3.0, importance_sample_ratio=0.75)
data = syn.training_batch(clips, H, W, T, device='cuda')
params = [p for p in det.parameters() if p.requires_grad]
opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.05)
def step():
    opt.zero_grad(set_to_none=True)
    out = det.train_step(data, opt)
    out['loss'].backward()
    torch.nn.utils.clip_grad_norm_(params, 0.01)
    opt.step()
    return out
for _ in range(2):
    o = step()
torch.cuda.synchronize()
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    o = step()
e1.record()
torch.cuda.synchronize()
print('clips', clips, 'ms/step', e0.elapsed_time(e1) / n, 'wall', (time.time() - t0) / n * 1e3, 'loss', float(o['loss']), 'mem GB', torch.cuda.max_memory_allocated() / 2**30)
if os.environ.get('PROFILE'):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=28, max_name_column_width=60))
if os.environ.get('CPROFILE'):
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
if os.environ.get('NCU'):
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
