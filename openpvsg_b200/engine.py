"""CUDA-graph frame runner for the VPS detector.

A frame's device work (backbone -> pixel decoder -> 9 decoder layers -> fused panoptic /
instance post-processing, ~350 kernel launches) has static shapes, no host round trip
and allocates only through torch's caching allocator, so it is captured once per input
shape into one CUDA graph and replayed: the launch-bound decoder stops paying Python +
launch latency.

Throughput mode: ``batch`` frames are pushed through the network together.  Frames stay
independent work items (clip length 1: the batch axis of every kernel is the frame axis), but
the latency-bound part of the path -- the decoder's 100-query chains, ~250 small launches per
pass -- is amortised over the batch and the tcgen05 GEMMs get more tiles per launch.

Per batch the host does: H2D copies into the static input buffer, one graph launch,
asynchronous D2H copies of the fixed-size outputs into a ring of pinned host buffers, and
an event record.  ``submit`` returns immediately; ``collect`` waits for that batch's event
and builds the reference's result dicts, so the host-side work of batch i overlaps the
device work of batch i+1 (``Mask2FormerVideoCustom.simple_test`` = submit + collect, batch 1).
"""
import os
import numpy as np
import torch

from . import lib as _l
from . import ops
from .mask2former import INSTANCE_OFFSET, bbox2result

RING = 3
SERIALISE_RUNNERS = os.environ.get('PVSG_SERIALISE_RUNNERS', '1') != '0'   # stream_frames: order the graphs of different runners
SYNC_CHUNKS = int(os.environ.get('PVSG_SYNC_CHUNKS', '2'))   # pieces a synchronous multi-sample call is pipelined in (simple_test)
DEBUG_MASKS = False   # parity tests: every runner also returns the decoder's sign masks and class logits
_copy_pool = None


def _pool():
    """Worker threads for the pinned -> caller-owned numpy copies of collect(copy=True) (np.copy releases
    the GIL; one thread moves ~10 GB/s, a batch of eight 720p results is 104 MB)."""
    global _copy_pool
    if _copy_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _copy_pool = ThreadPoolExecutor(max_workers=4)
    return _copy_pool
TOPK_INS = 10   # models/mask2former_vps/mask2former.py:192-195 keeps the 10 best instances


class _Pending:
    __slots__ = ('slot', 'event', 'n', 'gen')

    def __init__(self, slot, event, n, gen):
        self.slot, self.event, self.n, self.gen = slot, event, n, gen


def postprocess_batch(det, cls, mask_lr, in_hw, img_hw, out_hw):
    """Static-shape device post-processing of a batch of frames (CUDA-graph capturable), one launch
    per kernel for the whole batch: fused panoptic map + segment table, and the detector's top-10
    instances (models/mask2former_vps/mask2former.py:183-201): statistics of all max_per_image
    candidates first, binary masks only for the survivors.  cls [B,Q,NC+1], mask_lr [B,Q,h,w]."""
    fh = det.panoptic_fusion_head
    cfg = fh.test_cfg
    if not cfg.get('panoptic_on', True):
        raise NotImplementedError('FrameRunner needs panoptic_on (the VPS test configuration)')
    return ops.postprocess_batched(cls, mask_lr, in_hw, img_hw, out_hw, fh.num_things_classes, fh.num_classes,
                                   float(cfg.get('object_mask_thr', 0.8)), float(cfg.get('iou_thr', 0.8)),
                                   bool(cfg.get('filter_low_score', False)), INSTANCE_OFFSET,
                                   bool(cfg.get('instance_on', False)), cfg.get('max_per_image', 100), TOPK_INS)


class FrameRunner:
    """Captured per (H, W) frame shape for a ``Mask2FormerVideoCustom`` (clip length 1)."""

    def __init__(self, detector, meta, rescale=True, batch=1, lanes=None, rle=False, debug_masks=False):
        self.det = detector
        self.meta = dict(meta)
        self.rescale = rescale
        self.batch = int(batch)
        self.rle = bool(rle)      # also emit the tube wire format's run-length events (ops.rle_events)
        self.debug_masks = bool(debug_masks)   # parity tests: also return the decoder's sign masks (attn_mask_<layer>)
        dev = next(detector.parameters()).device
        self.dev = dev
        hp, wp = meta['batch_input_shape']
        # Lanes: independent instances of the captured graph (own input / activation / output
        # memory) replayed on their own streams, consecutive batches alternating between them, so
        # that one batch's latency-bound decoder chain could overlap the other's dense phases.
        # Measured on B200 (720p, batch 8): 2 lanes = 299 fps vs 311 fps with one -- the persistent
        # one-CTA-per-SM GEMMs of the two graphs just queue behind each other -- so the default is 1.
        self.nlanes = int(lanes) if lanes else 1
        self.lane_stream = [torch.cuda.Stream(device=dev) for _ in range(self.nlanes)]
        self.lane_in, self.lane_graph, self.lane_out = [], [], []
        self.launches_per_frame = 0
        for i in range(self.nlanes):
            self.static_in = torch.zeros(self.batch, 3, hp, wp, device=dev, dtype=torch.float32)
            self.graph = None
            self.out = None
            self._capture(first=i == 0)
            self.lane_in.append(self.static_in)
            self.lane_graph.append(self.graph)
            self.lane_out.append(self.out)
        # copy engines run beside the compute streams: frames go up on ``h2d``, results come down on
        # ``d2h`` (PCIe is full duplex), both through device-side staging slots so that batch i's
        # copies overlap batch i+1's graph
        self.h2d = torch.cuda.Stream(device=dev)
        self.d2h = torch.cuda.Stream(device=dev)
        self.in_stage = [torch.empty_like(self.static_in) for _ in range(RING)]
        self.in_ready = [torch.cuda.Event() for _ in range(RING)]
        self.in_free = [None] * RING
        self.out_stage = [{k: torch.empty_like(v) for k, v in self.out.items()} for _ in range(RING)]
        self.out_ready = [torch.cuda.Event() for _ in range(RING)]
        self.host = [{k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in self.out.items()}
                     for _ in range(RING)]
        self.events = [torch.cuda.Event() for _ in range(RING)]
        self.busy = [False] * RING
        self.gen = [0] * RING           # generation of the batch occupying a slot (stale-handle detection)
        self.pending = [None] * RING    # uncollected handle per slot
        self.next_slot = 0
        self.next_lane = 0

    @torch.no_grad()
    def _device_forward(self):
        det, meta, B = self.det, self.meta, self.batch
        feats = det.extract_feat(self.static_in)
        if self.debug_masks:
            det.panoptic_head._capture_masks = []
        try:
            cls, mask_lr, query = det.panoptic_head.simple_test_with_query(feats, [[meta]] * B, upsample=False)
        finally:
            captured, det.panoptic_head._capture_masks = det.panoptic_head._capture_masks, None
        fh = det.panoptic_fusion_head
        in_hw = tuple(meta['batch_input_shape'])
        img_hw = tuple(meta['img_shape'][:2])
        out_hw = tuple(meta['ori_shape'][:2]) if self.rescale else img_hw
        out = postprocess_batch(det, cls, mask_lr[:, 0].contiguous(), in_hw, img_hw, out_hw)
        out['query'] = query.transpose(0, 1).contiguous()      # [B,Q,C]
        if self.debug_masks:
            out['cls'] = cls.contiguous()
            for i, m in enumerate(captured):
                out[f'attn_mask_{i}'] = m
        if self.rle:
            out['rle_pos'], out['rle_slot'], out['rle_n'] = ops.rle_events(out['pan'], out['seg_info'])
        return out

    def _capture(self, first=True):
        _l.handle(torch.cuda.current_device())      # kernel attributes configured before anything is captured
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2 if first else 1):  # warm-up: fills weight / positional-encoding caches, no H2D left inside
                self._device_forward()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = _l.launch_count[0]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._device_forward()
        self.launches_per_frame = (_l.launch_count[0] - n0) / self.batch

    @torch.no_grad()
    def submit(self, imgs, after=None):
        """imgs: one frame ([1,3,H,W] / [3,H,W]) or a list of up to ``batch`` frames, device or
        pinned host tensors.  Enqueues H2D, the graph and the D2H of its outputs; returns a handle
        for ``collect``.  A short final batch is padded by repeating its last frame.
        after: optional CUDA event the graph replay waits for (the compute of ANOTHER runner: graphs of
        two runners live on different streams and would otherwise run concurrently, contending for the SMs)."""
        if torch.is_tensor(imgs):
            imgs = [imgs]
        n = len(imgs)
        if not 0 < n <= self.batch:
            raise ValueError(f'submit: expected 1..{self.batch} frames, got {n}')
        slot = self.next_slot
        if self.pending[slot] is not None:
            # the ring is full: this slot's pinned buffers and event still belong to an uncollected batch
            raise RuntimeError(f'FrameRunner.submit: {RING} batches are in flight; collect() the oldest one first')
        self.next_slot = (slot + 1) % RING
        lane = self.next_lane
        self.next_lane = (lane + 1) % self.nlanes
        caller = torch.cuda.current_stream()
        main = self.lane_stream[lane]
        static_in, out = self.lane_in[lane], self.lane_out[lane]
        shape = static_in.shape[1:]
        main.wait_stream(caller)          # device inputs were produced on the caller's stream
        with torch.cuda.stream(main):
            if any(not t.is_cuda for t in imgs):
                with torch.cuda.stream(self.h2d):
                    if self.in_free[slot] is not None:
                        self.h2d.wait_event(self.in_free[slot])      # the compute stream has drained this slot
                    for b in range(self.batch):
                        self.in_stage[slot][b].copy_(imgs[min(b, n - 1)].reshape(shape), non_blocking=True)
                    self.in_ready[slot].record(self.h2d)
                main.wait_event(self.in_ready[slot])
                static_in.copy_(self.in_stage[slot], non_blocking=True)
                self.in_free[slot] = torch.cuda.Event()
                self.in_free[slot].record(main)
            else:
                for b in range(self.batch):
                    static_in[b].copy_(imgs[min(b, n - 1)].reshape(shape), non_blocking=True)
            if after is not None:
                main.wait_event(after)
            self.lane_graph[lane].replay()
            self.last_compute = torch.cuda.Event()
            self.last_compute.record(main)
            if self.busy[slot]:
                main.wait_event(self.events[slot])                   # this slot's previous D2H has finished
            for k, v in out.items():
                self.out_stage[slot][k].copy_(v, non_blocking=True)
            self.out_ready[slot].record(main)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self.out_ready[slot])
            for k, v in self.out_stage[slot].items():
                self.host[slot][k].copy_(v, non_blocking=True)
            self.events[slot].record(self.d2h)
        self.busy[slot] = True
        self.gen[slot] += 1
        self.pending[slot] = _Pending(slot, self.events[slot], n, self.gen[slot])
        return self.pending[slot]

    @torch.no_grad()
    def collect(self, pending, copy=True):
        """Wait for a submitted batch and build the reference's per-frame result dicts
        (models/mask2former_vps/mask2former.py:172-211): a list with one dict per submitted frame.
        ``copy=False`` returns views of the pinned ring buffers (valid until RING-1 further submits)."""
        if self.pending[pending.slot] is not pending or pending.gen != self.gen[pending.slot]:
            raise RuntimeError('FrameRunner.collect: stale handle (already collected, or its ring slot was reused)')
        pending.event.synchronize()
        self.pending[pending.slot] = None   # the slot may be submitted again; copy=False views live until then
        det = self.det
        fh = det.panoptic_fusion_head
        own = (lambda a: a.copy()) if copy else (lambda a: a)
        big = {}
        if copy and pending.n > 1:      # the large arrays of all frames are copied concurrently
            hs = self.host[pending.slot]
            futs = {(k, b): _pool().submit(np.copy, hs[k][b].numpy()) for k in ('pan', 'ins_masks') if k in hs
                    for b in range(pending.n)}
            big = {kb: f.result() for kb, f in futs.items()}
        results = []
        for b in range(pending.n):
            hb = {k: v[b] for k, v in self.host[pending.slot].items()}
            res = {}
            if 'pan' in hb:
                res['pan_results'] = big[('pan', b)] if ('pan', b) in big else own(hb['pan'].numpy())
                query = hb['query'].clone() if copy else hb['query']
                res['query_feats'] = fh._query_dict(hb['seg_info'].numpy(), query)
                if 'rle_n' in hb and int(hb['rle_n']) <= hb['rle_pos'].numel():
                    from . import tubes
                    res['rle'] = tubes.rle_from_events(hb['rle_pos'].numpy(), hb['rle_slot'].numpy(), int(hb['rle_n']),
                                                       tubes.slot_ids(hb['seg_info'].numpy()), *hb['pan'].shape)
            if self.debug_masks:
                res['attn_masks'] = [hb[f'attn_mask_{i}'].numpy().copy() for i in range(len(hb)) if f'attn_mask_{i}' in hb]
                res['cls'] = hb['cls'].numpy().copy()
            if 'ins_boxes' in hb:
                n = min(TOPK_INS, int(hb['ins_count'][0]))
                labels = hb['ins_labels'][:n]
                bbox_results = bbox2result(hb['ins_boxes'][:n], labels, det.num_things_classes)
                masks_np = big[('ins_masks', b)][:n] if ('ins_masks', b) in big else hb['ins_masks'][:n].numpy()
                mask_results = [[] for _ in range(det.num_things_classes)]
                for j, label in enumerate(labels.tolist()):
                    mj = masks_np[j] if ('ins_masks', b) in big else own(masks_np[j])   # big[] is already caller-owned
                    mask_results[label].append(mj.view(np.bool_))
                res['ins_results'] = bbox_results, mask_results
            results.append(res)
        return results

    def run(self, img):
        return self.collect(self.submit(img))[0]


def enable_cuda_graph(detector):
    """Make ``Mask2FormerVideoCustom.simple_test`` replay a captured graph per frame shape."""
    detector._runners = {}
    detector._runners_epoch = weights_epoch(detector)
    return detector


def weights_epoch(detector):
    """Changes whenever a parameter / buffer is updated in place (load_state_dict, optimizer steps: the version
    counters) or the module is moved / cast / reloaded (``_DetectorBase._apply`` / ``_load_from_state_dict`` bump
    ``_weights_gen``).  A captured graph bakes in pointers to the kernel-layout copies of the weights (folded
    conv + BN tensors, bf16 operand planes), so it is only valid for the epoch it was captured in."""
    ts = detector.__dict__.get('_epoch_tensors')
    if ts is None or ts[0] != getattr(detector, '_weights_gen', 0):
        ts = (getattr(detector, '_weights_gen', 0), list(detector.parameters()) + list(detector.buffers()))
        detector.__dict__['_epoch_tensors'] = ts
    return hash((ts[0],) + tuple(t._version for t in ts[1]))


def get_runner(detector, meta, rescale=True, batch=1, rle=False, debug_masks=False):
    debug_masks = bool(debug_masks or DEBUG_MASKS)
    key = (tuple(meta['batch_input_shape']), tuple(meta['img_shape']), tuple(meta['ori_shape']), bool(rescale),
           int(batch), bool(rle), debug_masks)
    runners = detector._runners
    epoch = weights_epoch(detector)
    if getattr(detector, '_runners_epoch', None) != epoch:
        runners.clear()               # weights changed under the captured graphs: drop them, re-capture on demand
        detector._runners_epoch = epoch
    if key not in runners:
        runners[key] = FrameRunner(detector, meta, rescale, batch, rle=rle, debug_masks=debug_masks)
    return runners[key]


def batch_schedule(n, batch, ramp=True):
    """Batch sizes for n frames through runners of ``batch`` and ``batch // 2`` frames: a half batch first and last, so
    that the pipeline's fill (the first host->device copy, which nothing overlaps) and drain (the last device->host copy
    and result building) cost half a batch each.  n = 100, batch = 20 -> [10, 20, 20, 20, 20, 10]."""
    half = batch // 2
    if not ramp or half < 1 or n < 2 * batch:
        return [batch] * (n // batch) + ([n % batch] if n % batch else [])
    sched = [half]
    left = n - half
    while left > batch + half:
        sched.append(batch)
        left -= batch
    if left > batch:                      # batch < left <= batch + half
        sched += [left - half, half] if left - half <= batch else [batch, left - batch]
    else:
        sched.append(left)
    return [b for b in sched if b > 0]


@torch.no_grad()
def stream_frames(detector, meta, frames, batch, consume, rescale=True, rle=False, ramp=True):
    """Software-pipelined inference of a list of frames (pinned host or device tensors [3,H,W]) through the CUDA-graph
    runners: batch i+1 is submitted before batch i is collected; ``consume(result)`` is called per frame, in order, with
    views of the runner's pinned ring.  With ``ramp`` the first and last batch are half-sized (``batch_schedule``)."""
    if getattr(detector, '_runners', None) is None:
        enable_cuda_graph(detector)
    pend = None
    i = 0
    for b in batch_schedule(len(frames), batch, ramp):
        size = batch if b > batch // 2 else max(batch // 2, 1)        # the runner whose capacity fits this batch
        runner = get_runner(detector, meta, rescale, batch=size, rle=rle)
        after = pend[0].last_compute if (pend is not None and pend[0] is not runner and SERIALISE_RUNNERS) else None
        nxt = (runner, runner.submit(frames[i:i + b], after=after))
        i += b
        if pend is not None:
            for r in pend[0].collect(pend[1], copy=False):
                consume(r)
        pend = nxt
    if pend is not None:
        for r in pend[0].collect(pend[1], copy=False):
            consume(r)
