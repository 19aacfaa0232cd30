"""A/B of the e2e clip schedule (ramped vs plain batches) under torchrun: per-step time, max over ranks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import openpvsg_b200 as pv
from openpvsg_b200 import configs, engine, synthetic as syn
rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
det = pv.build_detector(configs.mask2former_r50(True))
det.load_state_dict(syn.mask2former_state_dict(seed=0))
det.cuda()
engine.enable_cuda_graph(det)
H, W = 720, 1280
meta = syn.frame_meta(H, W)
frames = [syn.synthetic_frame(s + 100 * rank, H, W).pin_memory() for s in range(20)]
clip = [frames[i % 20] for i in range(100)]
sink = []
def step(ramp):
    engine.stream_frames(det, meta, clip, 20, lambda r: sink.append(len(r['query_feats'])), ramp=ramp)
    sink.clear()
for ramp in (True, False):
    step(ramp)
res = {True: [], False: [], 'ramp, runners not ordered': []}
for rep in range(4):
    for ramp in (True, False, 'ramp, runners not ordered'):
        engine.SERIALISE_RUNNERS = ramp is not 'ramp, runners not ordered'
        ramp_flag = bool(ramp)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step(ramp_flag)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device='cuda')
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        res[ramp].append(float(dt) * 1e3)
if rank == 0:
    print('ramp  ', [round(x, 1) for x in res[True]])
    print('plain ', [round(x, 1) for x in res[False]])
    print('ramp, runners not ordered', [round(x, 1) for x in res['ramp, runners not ordered']])
if world > 1:
    dist.destroy_process_group()
