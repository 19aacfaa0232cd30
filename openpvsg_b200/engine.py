"""CUDA-graph frame runner for the VPS detector.

A frame's device work (backbone -> pixel decoder -> 9 decoder layers -> fused panoptic /
instance post-processing, ~450 kernel launches) has static shapes, no host round trip
and allocates only through torch's caching allocator, so it is captured once per input
shape into one CUDA graph and replayed per frame: the launch-bound decoder stops paying
Python + launch latency.  Host-side work per frame is one H2D copy into the static input
buffer, one graph launch, and the D2H copies of the results.
"""
import numpy as np
import torch

from . import lib as _l
from .mask2former import bbox2result


class FrameRunner:
    """Captured per (H, W) frame shape for a ``Mask2FormerVideoCustom`` (clip length 1)."""

    def __init__(self, detector, meta, rescale=True):
        self.det = detector
        self.meta = dict(meta)
        self.rescale = rescale
        dev = next(detector.parameters()).device
        hp, wp = meta['batch_input_shape']
        self.static_in = torch.zeros(1, 3, hp, wp, device=dev, dtype=torch.float32)
        self.graph = None
        self.out = None
        self.launches_per_frame = 0
        self._capture()

    @torch.no_grad()
    def _device_forward(self):
        det, meta = self.det, self.meta
        feats = det.extract_feat(self.static_in)
        cls, mask_lr, query = det.panoptic_head.simple_test_with_query(feats, [[meta]], upsample=False)
        fh = det.panoptic_fusion_head
        in_hw = tuple(meta['batch_input_shape'])
        img_hw = tuple(meta['img_shape'][:2])
        out_hw = tuple(meta['ori_shape'][:2]) if self.rescale else img_hw
        out = dict(cls=cls[0], mask_lr=mask_lr[0, 0], query=query[:, 0])
        if fh.test_cfg.get('panoptic_on', True):
            out['pan'], out['seg_info'] = fh._panoptic(cls[0], mask_lr[0, 0], in_hw, img_hw, out_hw)
        if fh.test_cfg.get('instance_on', False):
            out['ins'] = fh._instance_device(cls[0], mask_lr[0, 0], in_hw, img_hw, out_hw, True)
        return out

    def _capture(self):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):  # warm-up: fills weight / positional-encoding caches, no H2D left inside
                self._device_forward()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = _l.launch_count[0]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._device_forward()
        self.launches_per_frame = _l.launch_count[0] - n0

    @torch.no_grad()
    def run(self, img):
        """img [1,3,H,W] or [3,H,W], device or (pinned) host tensor.  Returns the static output
        tensors (valid until the next run)."""
        self.static_in.copy_(img.reshape(self.static_in.shape), non_blocking=True)
        self.graph.replay()
        return self.out

    @torch.no_grad()
    def results(self, out=None):
        """Static outputs -> the reference's per-frame result dict
        (models/mask2former_vps/mask2former.py:172-211)."""
        out = out or self.out
        det = self.det
        fh = det.panoptic_fusion_head
        res = {}
        if 'pan' in out:
            seg_info = out['seg_info'].cpu().numpy()          # sync point
            res['pan_results'] = out['pan'].cpu().numpy()
            res['query_feats'] = fh._query_dict(seg_info, out['query'].clone())
        if 'ins' in out:
            d = out['ins']
            is_thing = d['labels'] < det.num_things_classes
            stats = d['stats']
            det_scores = d['scores'] * stats[:, 0] / (stats[:, 1] + 1e-6)
            det_scores = torch.where(is_thing, det_scores, det_scores.new_full((), -1.0))
            # ids are 1-based ranks among the thing candidates, as torch.arange(len(bboxes)) + 1 (:188)
            ids = torch.cumsum(is_thing.to(torch.float32), 0)
            n_thing = int(is_thing.sum().item())
            inds = torch.argsort(det_scores, descending=True)[:min(10, n_thing)]
            bboxes = torch.cat([ids[inds, None], d['boxes'][inds].float(), det_scores[inds, None]], dim=1)
            labels = d['labels'][inds]
            masks_np = d['masks'][inds].cpu().numpy().astype(bool)
            bbox_results = bbox2result(bboxes, labels, det.num_things_classes)
            mask_results = [[] for _ in range(det.num_things_classes)]
            for j, label in enumerate(labels.tolist()):
                mask_results[label].append(masks_np[j])
            res['ins_results'] = bbox_results, mask_results
        return res


def enable_cuda_graph(detector):
    """Make ``Mask2FormerVideoCustom.simple_test`` replay a captured graph per frame shape."""
    detector._runners = {}
    return detector


def get_runner(detector, meta, rescale=True):
    key = (tuple(meta['batch_input_shape']), tuple(meta['img_shape']), tuple(meta['ori_shape']), bool(rescale))
    runners = detector._runners
    if key not in runners:
        runners[key] = FrameRunner(detector, meta, rescale)
    return runners[key]
