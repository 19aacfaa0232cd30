"""Debug: ReLU-mask consistency of the FFN-up layers in the training forward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from openpvsg_b200 import ops, train_ops as T
import test_training_slice as tt

det, sd, frames, metas = tt._train_setup()
head = det.panoptic_head
rec = []
orig = T.linear
def spy(x, weight, bias=None, add_input=None, residual=None, act=ops.ACT_NONE):
    y = orig(x, weight, bias, add_input, residual, act)
    if act == ops.ACT_RELU and weight.shape[0] == 2048:
        rec.append((x.detach(), weight.detach(), bias.detach(), y.detach()))
    return y
T.linear = spy
with torch.no_grad():
    feats = det.extract_feat(frames[0].cuda())
cls_list, mask_list = head.forward_train_outputs(feats, 2)
for i, (x, w, b, y) in enumerate(rec):
    pre = x.double().reshape(-1, 256) @ w.double().T + b.double()
    flips = (pre > 0) != (y.reshape(-1, 2048) > 0)
    err = (torch.relu(pre) - y.reshape(-1, 2048).double()).abs().max().item()
    print(i, 'rows', pre.shape[0], 'flips', int(flips.sum()), 'max |pre| at flips', float(pre[flips].abs().max()) if flips.any() else 0.0,
          'max fwd err', err, 'frac active', float((y > 0).float().mean()))
