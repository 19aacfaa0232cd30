"""Throughput of the reference's synchronous call model(return_loss=False, ...) with 20 samples per call, for different
numbers of pipeline pieces inside the call (engine.SYNC_CHUNKS)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import openpvsg_b200 as pv
from openpvsg_b200 import configs, engine, synthetic as syn
det = pv.build_detector(configs.mask2former_r50(True))
det.load_state_dict(syn.mask2former_state_dict(seed=0))
det.cuda()
engine.enable_cuda_graph(det)
H, W, B = 720, 1280, 20
frames = [syn.synthetic_frame(s, H, W).pin_memory() for s in range(B)]
ref = torch.stack(frames)[:, None].pin_memory()
meta = syn.frame_meta(H, W)
for chunks in (1, 2, 4, 5):
    engine.SYNC_CHUNKS = chunks
    call = lambda: det(return_loss=False, rescale=True, img=[ref[:, 0]], img_metas=[[dict(meta) for _ in range(B)]], ref_img=[ref],
                       ref_img_metas=[[dict(meta)] for _ in range(B)])
    for _ in range(3):
        out = call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        out = call()
    dt = (time.perf_counter() - t0) / n
    print(f'chunks {chunks}: {dt * 1e3:.1f} ms per call of {B} = {B / dt:.1f} frames/s, results {len(out)}')
