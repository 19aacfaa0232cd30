"""Times the Swin-B backbone / detector stages at 720p (eager, CUDA events).  usage: python tools/swin_time.py [batch]"""
import sys
import torch
sys.path.insert(0, '.')
import openpvsg_b200 as pv
from openpvsg_b200 import configs, synthetic as syn

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device('cuda')
sd = syn.mask2former_state_dict(seed=5, in_channels=(128, 256, 512, 1024), backbone=dict(configs.SWIN_B))
det = pv.build_detector(configs.mask2former_swin(True))
det.load_state_dict(sd, strict=True)
det = det.to(dev)
x = torch.stack([syn.synthetic_frame(i, 720, 1280) for i in range(B)]).to(dev)


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, out


t_bb, feats = timed(lambda: det.backbone(x))
print(f'swin-b backbone: {t_bb:.2f} ms for {B} frames = {t_bb / B:.2f} ms/frame')
t_pd, _ = timed(lambda: det.panoptic_head.pixel_decoder(feats))
print(f'pixel decoder:   {t_pd:.2f} ms = {t_pd / B:.2f} ms/frame')
from openpvsg_b200 import ops
bb = det.backbone
tok = torch.randn(B, 184, 320, 128, device=dev)
blk = bb.stages[0].blocks[1]
w = blk.attn.w_msa
qkv = ops.linear(tok, w.qkv.weight, w.qkv.bias)
for st, (h, wd) in enumerate(((184, 320), (92, 160), (46, 80), (23, 40))):
    C = 128 << st
    m = bb.stages[st].blocks[1].attn
    q = torch.randn(B, h, wd, 3 * C, device=dev)
    t, _ = timed(lambda: ops.window_attention(q, m.w_msa.qkv.bias, m.w_msa.relative_position_bias_table, m.num_heads, 12, 6))
    print(f'window attention stage {st}: {t * 1e3:.0f} us for {B} frames')
    xx = torch.randn(B, h, wd, C, device=dev)
    t, _ = timed(lambda: bb.stages[st].blocks[1].forward_tokens(xx))
    print(f'  whole block: {t * 1e3:.0f} us')
