#!/usr/bin/env python
"""One replay of the 720p frame graph (default: the 20 frames per replay bench.py uses) between cudaProfilerStart/Stop (run under
`ncu --profile-from-start off`): the launch list of exactly the kernels bench.py times."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openpvsg_b200 as pv  # noqa: E402
from openpvsg_b200 import configs, engine, synthetic as syn  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 20
swin = len(sys.argv) > 2 and sys.argv[2] == 'swin_b'      # python tools/ncu_frame.py 8 swin_b
dev = torch.device('cuda:0')
if swin:
    det = pv.build_detector(configs.mask2former_swin(True))
    det.load_state_dict(syn.mask2former_state_dict(seed=0, in_channels=(128, 256, 512, 1024), mask_shift=32.0,
                                                   backbone=dict(configs.SWIN_B)))
else:
    det = pv.build_detector(configs.mask2former_r50(True))
    det.load_state_dict(syn.mask2former_state_dict(seed=0))
det.to(dev)
engine.enable_cuda_graph(det)
meta = syn.frame_meta(720, 1280)
frames = [syn.synthetic_frame(i, 720, 1280).to(dev) for i in range(batch)]
runner = engine.get_runner(det, meta, True, batch=batch)
runner.collect(runner.submit(frames))
torch.cuda.synchronize()
torch.cuda.profiler.start()
runner.collect(runner.submit(frames))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('launches per frame', runner.launches_per_frame)
