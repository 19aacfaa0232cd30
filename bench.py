#!/usr/bin/env python
"""Headline benchmark: Mask2Former-VPS (R50) inference frames/s on synthetic 720p clips.

  python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

A step = one pass of the hot path over one clip shard: `--frames` (default 100) synthetic
720p frames per GPU (BASELINE.json configs[1]; weak scaling: the clip has N x 100 frames,
contiguous blocks per rank), each frame through backbone -> pixel decoder (6 x MSDeformAttn)
-> 9 masked-attention decoder layers -> fused panoptic + instance post-processing, then the
tube-linking exchange (all-gather over NCCL when N > 1).

`value`  : frames/s with the frames already resident in HBM (results still read back).
`e2e`    : frames/s through the public detector API (model(return_loss=False, rescale=True,
           img=..., ref_img=...)) from PINNED HOST frames: H2D of every frame and D2H of every
           result inside the timed region.
`roofline`: the dominant kernel family (the fp32 GEMM / implicit-GEMM conv engine), achieved
           TFLOP/s on algorithmic FLOPs measured live with CUDA events on an instrumented
           frame, against the measured bf16 tensor peak of MEASURED_PEAKS.json.
`cpu_baseline`: the CPU oracle (a port of the reference algorithm; the reference itself needs
           mmcv/mmdet, which cannot be installed offline) timed on the host cores on a bounded
           sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H, W = 720, 1280
METRIC = 'Mask2Former-VPS R50 inference frames/sec @720p'
MSDA_KERNEL = 'msda_tile_kernel (csrc/msda_tile.cu: TMA-staged value windows in shared memory)'
MSDA_NOTE = ('bound by the LSU data pipe (shared-memory crossbar, 128 B/clk/SM), not by HBM: 12 samples x 4 corners x 128 B per '
             '(query, head) = 48 wavefronts minimum = 27 us per 720p frame and layer at 100 % of that pipe, against 9.5 us for the '
             'algorithmic HBM bytes, i.e. a ceiling of ~0.35 of the HBM roofline for fp32 values; ncu: data pipe 75 % busy, DRAM traffic '
             '0.99 x algorithmic (profiles/r02c_ncu_msda_tile.json, r02c_msda_traffic.json)')


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d['bf16_tflops_sustained'],
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback (B200_PROFILING.md)')


def kernel_traffic(family):
    """DRAM bytes per launch of a kernel family, from the committed ncu pass over one frame-graph replay
    (profiles/*_<family>_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum, averaged over launches)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', f'*_{family}_traffic.json')))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    return d.get('avg_bytes_per_launch'), os.path.relpath(files[-1], ROOT)


DTYPE = 'f32 io / bf16x3 mma (split-bf16 operands, ~16-bit mantissa per product) / f32 accumulate'


def workload_config(args, world):
    """The `config` object of the JSON line -- identical for both arms (the reference arm runs a bounded sample
    of the same workload; what it sampled is in its `cpu_baseline.sample`)."""
    swin = args.backbone == 'swin_b'
    workload = ('Mask2Former-VPS Swin-B inference (mmdet 2.25 Swin-B: embed 128, depths 2-2-18-2, window 12), synthetic 720p '
                'clip, 100 frames per GPU (the backbone of BASELINE configs[2]); random-init weights'
                if swin else
                'Mask2Former-VPS R50 inference, synthetic 720p clip, 100 frames per GPU '
                '(BASELINE configs[1]); random-init weights of the reference architecture')
    frames, distinct = getattr(args, 'frames', 100), getattr(args, 'distinct', 100)
    return dict(workload=workload, backbone=args.backbone, frames_per_gpu_per_step=frames,
                distinct_frames=distinct, frame_seeds=f'rank * {distinct} + 0..{distinct - 1}',
                resolution='720x1280 padded to 736x1280', clip_length=1, parallelism=f'frames x{world}',
                l2='per-frame working set (~1.5 GB of activations) >> 126 MB L2, no explicit flush',
                cuda_graph=not getattr(args, 'no_graph', False), frames_per_launch=getattr(args, 'batch', 20))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=smax, reasons=sorted(reasons),
                    samples=len(sm))


def make_frames(n_distinct, rank):
    """100 distinct frames, seeds 0..99 (rank r of a weak-scaled run owns seeds 100 r .. 100 r + 99)."""
    from concurrent.futures import ThreadPoolExecutor
    from openpvsg_b200 import synthetic as syn
    with ThreadPoolExecutor(8) as ex:
        return list(ex.map(lambda i: syn.synthetic_frame(n_distinct * rank + i, H, W), range(n_distinct)))


# --------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle on the host cores
# --------------------------------------------------------------------------------------
# mask_shift: the synthetic head is calibrated per backbone so that the fusion head keeps a realistic set of segments
# (Swin-B features with the R50 value give half-plane masks and an all-void panoptic map; 32 -> ~10 segments per frame)
SWIN_SD = dict(in_channels=(128, 256, 512, 1024), mask_shift=32.0,
               backbone=dict(embed_dims=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32), window_size=12))


def cpu_oracle_fps(n_frames, warmup=1, swin=False):
    from openpvsg_b200 import synthetic as syn
    from oracle import m2f as om
    torch.set_num_threads(os.cpu_count())
    sd = syn.mask2former_state_dict(seed=0, **SWIN_SD) if swin else syn.mask2former_state_dict(seed=0)
    kw = {}
    if swin:
        from oracle import swin as osw
        kw['backbone'] = lambda sd_, x: osw.swin_forward(sd_, x, prefix='backbone.', **SWIN_SD['backbone'])
    meta = syn.frame_meta(H, W)
    frames = [syn.synthetic_frame(i, H, W) for i in range(max(1, min(n_frames, 2)))]
    times = []
    with torch.no_grad():
        for i in range(warmup + n_frames):
            t0 = time.perf_counter()
            om.vps_simple_test(sd, frames[i % len(frames)][None, None], [[meta]], rescale=True, **kw)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return n_frames / sum(times), times


def relation_bench(dev, with_cpu):
    """BASELINE configs[3]: relation_head pair-transformer forward on 200 synthetic query tubes x 128
    frames (tools/rel_test.py:39-67), device time per forward; CPU oracle (the reference's own torch
    modules restated in oracle/relation.py) once on the host cores."""
    from openpvsg_b200 import relation_head as rh, synthetic as syn
    sds = syn.relation_state_dicts(seed=1)
    feats = torch.randn(200, 128, 256, generator=torch.Generator().manual_seed(0))
    mods = [rh.ObjectEncoder(256), rh.ObjectEncoder(256), rh.PairProposalNetwork(256, 1024), rh.TemporalTransformer(512, 57)]
    for m, k in zip(mods, ('subject_encoder', 'object_encoder', 'pair_proposal_model', 'relation_model')):
        m.load_state_dict(sds[k])
        m.to(dev)
    fd = feats.to(dev)
    for _ in range(3):
        rh.relation_forward(*mods, fd, 100)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rh.relation_forward(*mods, fd, 100)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out = dict(workload='relation head forward, 200 tubes x 128 frames, top-100 pairs (BASELINE configs[3])',
               ms=round(float(np.median(ts)), 3), forwards_per_s=round(1e3 / float(np.median(ts)), 1))
    for _ in range(3):
        rh.relation_forward(*mods, fd, 100, graph=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rh.relation_forward(*mods, fd, 100, graph=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out['ms_cuda_graph'] = round(float(np.median(ts)), 3)
    out['gflop'] = round(2 * 64.2 + 0.33 + 48.2, 1)
    out['tflops_cuda_graph'] = round(out['gflop'] / out['ms_cuda_graph'], 1)
    if with_cpu:
        from oracle import relation as orel
        torch.set_num_threads(os.cpu_count())
        with torch.no_grad():
            t0 = time.perf_counter()
            orel.relation_forward(sds, feats, 100)
            out['cpu_ms'] = round(1e3 * (time.perf_counter() - t0), 1)
        out['cpu_cores'] = os.cpu_count()
    return out


def relset_bench(dev, with_cpu, frames=20, num_gt=32, slots=100):
    """SURVEY 8f rank 1 (relation-set builder): per-frame overlap of a GT instance-id map with the panoptic map
    (``pvsg_tube_overlap``) against the HBM roofline; algorithmic bytes = two int32 maps per frame.  CPU baseline =
    the reference's evaluation (utils/relation_matching.py:156-165,205-260): one logical_and + logical_or pass per
    (GT object, same-class tube) pair of a frame, here for 12 objects x 4 candidate tubes on one frame."""
    from openpvsg_b200 import ops
    g = torch.Generator().manual_seed(0)
    # blocky label maps with ~1.4e4 (gt, segment) runs per frame -- the run density of real 720p panoptic maps
    gt = torch.randint(0, num_gt, (frames, H // 80, W // 80), generator=g).repeat_interleave(80, 1).repeat_interleave(80, 2).int()
    slot = torch.randint(0, 40, (frames, H // 16, W // 64), generator=g).repeat_interleave(16, 1).repeat_interleave(64, 2)
    ids = (torch.arange(slots) % 127 + 1000 * (torch.arange(slots) // 3)).int()
    pan = ids[slot].int()
    seg_info = torch.zeros(frames, 1 + 4 * slots, dtype=torch.int32)
    seg_info[:, 0] = 40
    seg_info[:, 3:3 + 4 * 40:4] = ids[:40]
    gt_d, pan_d, si_d = gt.to(dev), pan.to(dev), seg_info.to(dev)
    for _ in range(3):
        ops.tube_overlap(gt_d, pan_d, si_d, num_gt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        counts = ops.tube_overlap(gt_d, pan_d, si_d, num_gt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gbytes = 2 * 4 * frames * H * W / 1e9
    out = dict(frames=frames, ms=round(ms, 4), us_per_frame=round(1e3 * ms / frames, 2), achieved_GBps=round(gbytes / (ms * 1e-3), 1),
               algorithmic_bytes_per_frame=2 * 4 * H * W, checksum=int(counts.sum().item()),
               note='inputs (2 x 74 MB) fit the 126 MB L2 only partly; 20 back-to-back launches over the same maps')
    if with_cpu:
        gm, pm = gt[0].numpy(), pan[0].numpy()
        t0 = time.perf_counter()
        hits = 0
        for obj in range(12):
            gmask = gm == obj
            for k in range(4):
                pmask = pm == int(ids[(obj + k) % 40])
                inter, union = np.logical_and(gmask, pmask).sum(), np.logical_or(gmask, pmask).sum()
                hits += int(union > 0 and inter / union > 0.5)
        out['cpu_ms_per_frame'] = round(1e3 * (time.perf_counter() - t0), 2)
        out['cpu_sample'] = '1 frame, 12 GT objects x 4 candidate tubes, numpy masks (RLE decode not included)'
    return out


def tube_dump_bench(det, meta, frames, batch):
    """masks.txt rows of a batch of frames -- device run-length events
    (pvsg_rle_events + tubes.rle_from_events) vs the host encoder working on the panoptic map
    (what concat_seq does with pycocotools in the reference)."""
    from openpvsg_b200 import engine, tubes
    runner = engine.get_runner(det, meta, True, batch=batch, rle=True)
    pend = runner.submit(frames[:batch])
    res = runner.collect(pend)
    hb = runner.host[pend.slot]
    H_, W_ = hb['pan'].shape[1:]
    t0 = time.perf_counter()
    dev_rows = 0
    for b in range(batch):      # what collect() does per frame: events -> strings (native host routine)
        ids = tubes.slot_ids(hb['seg_info'][b].numpy())
        dev_rows += len(tubes.rle_from_events(hb['rle_pos'][b].numpy(), hb['rle_slot'][b].numpy(), int(hb['rle_n'][b]),
                                              ids, H_, W_))
    t_dev = (time.perf_counter() - t0) / batch
    t0 = time.perf_counter()
    host_rows = 0
    for r in res:
        for sid in r['query_feats']:
            tubes.rle_string(tubes.rle_counts(r['pan_results'] == sid))
            host_rows += 1
    t_host = (time.perf_counter() - t0) / batch
    return dict(workload='masks.txt RLE rows per 720p frame (tube wire format)', segments_per_frame=round(host_rows / batch, 1),
                events_per_frame=int(hb['rle_n'][:batch].float().mean()),
                device_events_host_ms_per_frame=round(1e3 * t_dev, 2), host_encoder_ms_per_frame=round(1e3 * t_host, 2),
                note='device path: pvsg_rle_events inside the frame graph (3 launches per batch) + pvsg_rle_strings_host; '
                     'host path: numpy RLE of pan == id per segment (the reference uses pycocotools per segment)')


def training_bench(dev, with_cpu, clips=16, steps=3, world=1, rank=0):
    """one optimisation step of the VPS detector at the reference's training configuration
    (configs/mask2former_vps/mask2former_video_r50.py: samples_per_gpu = 16 clips of 2 frames, 360 x 480 padded to
    384 x 480, 12544 loss points; _base_/schedules/m2f_schedules.py: AdamW lr 1e-4, weight decay 0.05, gradient clipping at
    0.01): detector.train_step (forward_train + the 30 loss terms) -> backward -> clip -> AdamW step, every forward and
    backward op a library kernel; device time per step.  CPU: the oracle's forward + torch autograd backward of ONE clip."""
    import torch.distributed as dist
    import openpvsg_b200 as pv
    from openpvsg_b200 import configs, synthetic as syn
    det = pv.build_detector(configs.mask2former_r50(True))
    det.load_state_dict(syn.mask2former_state_dict(seed=3))
    det.to(dev)
    det.panoptic_head.train_cfg = dict(num_points=12544, oversample_ratio=3.0, importance_sample_ratio=0.75)
    from openpvsg_b200 import dist_train
    data = syn.training_batch(clips, device=dev, seed=1000 * rank)      # every rank trains on its own clips (weak scaling)
    params = [p for p in det.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.05)
    dist_train.broadcast_parameters(det)
    state = dict(bucket=None)

    def step():
        opt.zero_grad(set_to_none=True)
        out = det.train_step(data, opt)
        out['loss'].backward()
        state['bucket'] = dist_train.allreduce_gradients(params, state['bucket'])   # the one exchange of data-parallel training
        torch.nn.utils.clip_grad_norm_(params, 0.01)
        opt.step()
        return out

    from openpvsg_b200 import lib as _l
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = _l.launch_count[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.reset_peak_memory_stats()
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:                                   # max over ranks
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clips_rank, clips = clips, clips * world
    res = dict(workload=f'VPS training step: {world} GPU(s) x {clips_rank} clips x 2 frames @384x480 (reference training config), forward_train + '
                        'backward + grad clip + AdamW, every parameter trainable (BatchNorm in eval mode with '
                        'trainable affine, as the reference config)', ms_per_step=round(ms, 1), clips_per_s=round(clips / ms * 1e3, 2),
               frames_per_s=round(2 * clips / ms * 1e3, 2), trainable_tensors=len(params), loss=round(float(out['loss'].detach()), 3),
               scaling='weak: clips sharded over the ranks, one flat gradient all-reduce (176 MB) per step',
               library_calls_per_step=int((_l.launch_count[0] - n0) / steps),
               peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 1))
    del det, opt, params, data, out
    torch.cuda.empty_cache()
    if with_cpu and rank == 0:
        from oracle import losses as ol, m2f as om
        torch.set_num_threads(os.cpu_count())
        sd = {k: v.clone().float() for k, v in syn.mask2former_state_dict(seed=3).items()}
        for k, v in sd.items():
            if v.is_floating_point() and 'running_' not in k:
                v.requires_grad_(True)
        one = syn.training_batch(1)
        gt = torch.stack(one['ref_gt_masks'][0], 1).float()                       # [G,T,H,W]
        labels = one['ref_gt_labels'][0][:gt.shape[0], 1]
        g = torch.Generator().manual_seed(0)
        t0 = time.perf_counter()
        cls, masks, _ = om.head_forward(sd, om.resnet50(sd, one['ref_img'][0]), video=True, num_frames=2)
        total = 0
        for c, m in zip(cls, masks):
            lc, lm, ld, _, _ = ol.loss_single(c, m, [labels], [gt], torch.rand(1, 12544, 2, generator=g),
                                              lambda n: torch.rand(n, 12544, 2, generator=g))
            total = total + lc + lm + ld
        total.backward()
        sec = time.perf_counter() - t0
        res.update(cpu_clips_per_s=round(1.0 / sec, 4), cpu_cores=os.cpu_count(),
                   cpu_sample='1 clip of 2 frames through the CPU oracle forward + torch autograd backward (no optimizer step)')
    return res


def _relation_models(dev):
    from openpvsg_b200 import relation_head as rh, synthetic as syn
    sds = syn.relation_state_dicts(seed=1)
    mods = [rh.ObjectEncoder(256), rh.ObjectEncoder(256), rh.PairProposalNetwork(256, 1024), rh.TemporalTransformer(512, 57)]
    for m, k in zip(mods, ('subject_encoder', 'object_encoder', 'pair_proposal_model', 'relation_model')):
        m.load_state_dict(sds[k])
        m.to(dev)
    return mods, sds


def end2end_bench(det, meta, host_frames, batch, dev, world, rank, timed, with_cpu=True):
    """BASELINE configs[4]: 380 frames (76 s @ 5 FPS) @720p through VPS -> tube linking (device RLE rows; all-gather of
    the kept entries when the clip is sharded over `world` ranks) -> relation head, from pinned host frames, timed
    between barriers as the max over ranks.  With the CPU baseline enabled, rank 0 also runs the CPU oracle's relation
    stage on the same tubes and reports R@20/50/100 of both backends against synthetic ground truth drawn from the
    oracle's own top triplets (the two must be equal)."""
    from openpvsg_b200 import end2end, rel_eval, tubes
    mods, sds = _relation_models(dev)
    T = 380
    lo, hi = tubes.shard_frames(T, world, rank)
    clip = [host_frames[i % len(host_frames)] for i in range(lo, hi)]   # this rank's contiguous block

    def run():
        return end2end.run_clip(det, mods, clip, meta, batch=batch, num_frames=T)

    ms, out, _ = timed(run, 1, 1)          # warm-up run captures the RLE graph / relation kernels
    res = dict(workload=f'end-to-end VPS + tube linking + relation head, 380 frames @720p (BASELINE configs[4]), {world} GPU(s), '
                        'frames sharded in contiguous blocks, one all-gather at tube linking',
               seconds=round(ms * 1e-3, 3), frames_per_s=round(T / (ms * 1e-3), 1), tubes=len(out['linker'].object_list),
               mask_rows_rank0=len(out['linker'].rows), triplets=len(out['relations']))
    if with_cpu and rank == 0 and out['raw'] is not None:
        from oracle import relation as orel        # checker only (cpu_baseline leg)
        feats = torch.as_tensor(out['linker'].tube_features())
        torch.set_num_threads(os.cpu_count())
        with torch.no_grad():
            ref = orel.relation_forward(sds, feats, 100)
        ref_res = orel.generate_pairwise_results(ref['span_pred'], ref['prob'], ref['pairs'])
        gt = [dict(subject_index=r['subject_index'], object_index=r['object_index'], relation=r['relation'],
                   relation_span=np.asarray(r['relation_span'])) for r in ref_res[::3][:30]]
        from openpvsg_b200 import relation_head as rh
        gpu_res = rh.generate_pairwise_results(out['raw']['span_pred'], out['raw']['prob'], out['raw']['pairs'].cpu().tolist())
        names = [f'relation_{i}' for i in range(57)]
        recalls = {}
        for tag, results in (('b200', gpu_res), ('cpu_oracle', ref_res)):
            d = rel_eval.new_recall_dict(names)
            rel_eval.accumulate(d, results, gt)
            fm = rel_eval.calculate_final_metrics(d, list(rel_eval.K_VALUES))
            recalls[tag] = {f'R@{K}': round(fm[K]['recall'], 6) for K in rel_eval.K_VALUES}
        res['recall'] = dict(recalls, equal=recalls['b200'] == recalls['cpu_oracle'], gt_relations=len(gt),
                             note='relation stage of the CPU oracle on the same tube features; GT = every 3rd of the '
                                  "oracle's top pairwise triplets (first 30)")
    return res


def swin_clip_bench(pv, configs, engine, syn, tubes, meta, host_frames, batch, dev, world, rank, timed):
    """BASELINE configs[2]: Mask2Former-VPS with the Swin-B backbone on a 300-frame 720p clip sharded over the ranks
    (contiguous blocks of ceil(300 / N) frames), pinned host frames in, results + tube linking out."""
    det = pv.build_detector(configs.mask2former_swin(True))
    det.load_state_dict(syn.mask2former_state_dict(seed=0, **SWIN_SD))
    det.to(dev)
    engine.enable_cuda_graph(det)
    T = 300
    lo, hi = tubes.shard_frames(T, world, rank)
    clip = [host_frames[i % len(host_frames)] for i in range(lo, hi)]
    runner = engine.get_runner(det, meta, True, batch=batch)

    def run():
        entries, pend = [], None

        def consume(res):
            ids = list(res['query_feats'].keys())
            entries.append((ids, np.stack([np.asarray(res['query_feats'][k][0]) for k in ids]) if ids
                            else np.zeros((0, 256), np.float32)))

        for i in range(0, len(clip), batch):
            nxt = runner.submit(clip[i:i + batch])
            if pend is not None:
                for r in runner.collect(pend, copy=False):
                    consume(r)
            pend = nxt
        for r in runner.collect(pend, copy=False):
            consume(r)
        return tubes.gather_and_link(entries, T, device=dev)

    ms, linker, _ = timed(run, 2, 1)
    out = dict(workload=f'Mask2Former-VPS Swin-B inference, synthetic 720p 300-frame clip sharded across {world} GPU(s) '
                        '(BASELINE configs[2]; Swin-B is not in the reference: oracle-checked by us, SURVEY 8d)',
               seconds_per_clip=round(ms * 1e-3 / 2, 3), frames_per_s=round(2 * T / (ms * 1e-3), 1), tubes=len(linker.object_list),
               frames_per_rank=hi - lo, api='engine.FrameRunner.submit/collect on pinned host frames + tubes.gather_and_link')
    det._runners.clear()
    del runner, det
    torch.cuda.empty_cache()
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    swin = args.backbone == 'swin_b'
    fps, times = cpu_oracle_fps(args.steps, warmup=min(args.warmup, 1), swin=swin)
    cores = os.cpu_count()
    line = dict(metric=METRIC.replace('R50', 'Swin-B') if swin else METRIC, value=fps, unit='frames/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * float(np.mean(times)), higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic', impl='reference',
                config=workload_config(args, args.gpus),
                note='CPU oracle port of the reference algorithm (mmcv / mmdet are not installable offline, so the reference '
                     'itself cannot run); one step = ONE frame of the workload (bounded sample)',
                cpu_baseline=dict(value=fps, unit='frames/s', cores=cores, kind='port',
                                  sample=f'{args.steps} frame(s) @720p, 1 frame per step, torch CPU fp32, {cores} threads'),
                e2e=dict(value=fps, unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# instrumented frame: per-kernel-family time / flops with CUDA events
# --------------------------------------------------------------------------------------
def kernel_breakdown(det, img, meta, batch=1):
    """Device time per kernel family for one frame.

    One eager frame is run with every ops.* call recorded (function, arguments).  Then, per
    family, the recorded calls are replayed inside ONE CUDA graph (same input tensors, outputs
    discarded) and the graph is timed with CUDA events: pure device time, no host launch latency,
    same kernels / shapes / launch order as the timed region."""
    from openpvsg_b200 import ops
    calls = []

    def lin_flops(a, k):
        x, w = a[0], a[1]
        return 2.0 * (int(np.prod(x.shape[:-1]))) * w.shape[0] * w.shape[1]

    def conv_flops(a, k):
        x, w = a[0], a[1]
        s = k.get('stride', 1)
        return 2.0 * x.shape[0] * (x.shape[1] // s) * (x.shape[2] // s) * w.shape[0] * w.shape[1] * w.shape[2] * w.shape[3]

    def ml_flops(a, k):
        e, f = a[0], a[1]
        return 2.0 * e.shape[0] * e.shape[1] * e.shape[2] * f.shape[1]

    def msda_bytes(a, k):
        v, proj = a[0], a[2]
        return 4.0 * (2 * v.numel() + proj.numel())

    def numel(t):
        return int(np.prod(t.shape))

    def lin_bytes(a, k):      # fp32-equivalent operand + result bytes (a plane pair is 2 + 2 bytes per element)
        x, w = a[0], a[1]
        mn = numel(x) // x.shape[-1] * w.shape[0]
        res = k.get('residual', a[4] if len(a) > 4 else None)
        return 4.0 * (numel(x) + numel(w) + mn + (mn if res is not None else 0))

    def conv_bytes(a, k):
        x, w = a[0], a[1]
        s = k.get('stride', 1)
        out = x.shape[0] * (x.shape[1] // s) * (x.shape[2] // s) * w.shape[0]
        return 4.0 * (numel(x) + numel(w) + out + (out if k.get('residual') is not None else 0))

    def ml_bytes(a, k):
        e, f = a[0], a[1]
        want_logits = a[2] if len(a) > 2 else k.get('want_logits', True)
        want_mask = a[3] if len(a) > 3 else k.get('want_mask', False)
        qp = e.shape[0] * e.shape[1] * f.shape[1]
        return 4.0 * (numel(e) + numel(f)) + (4.0 * qp if want_logits else 0.0) + (1.0 * qp if want_mask else 0.0)

    spec = (('linear', 'gemm', lin_flops, lin_bytes), ('conv2d_nhwc', 'gemm', conv_flops, conv_bytes),
            ('mask_logits', 'gemm', ml_flops, ml_bytes), ('split_bf16', 'gemm', None, None),
            ('msda_fused_forward', 'msda', None, msda_bytes), ('attention', 'attention', None, None),
            ('layernorm', 'norm', None, None), ('groupnorm_nhwc', 'norm', None, None),
            ('add_rowvec', 'norm', None, None), ('panoptic_fuse', 'postprocess', None, None),
            ('instance_masks', 'postprocess', None, None), ('instance_select', 'postprocess', None, None),
            ('postprocess_batched', 'postprocess', None, None), ('bilinear_resize_nhwc', 'resize', None, None),
            ('maxpool3x3s2_nhwc', 'resize', None, None), ('window_attention', 'window_attention', None, None),
            ('patch_merge_ln', 'norm', None, None))
    saved, depth = {}, [0]

    def wrap(name, family, flops_fn, bytes_fn):
        orig = getattr(ops, name)
        saved[name] = orig

        def f(*a, **k):
            top = depth[0] == 0          # nested calls (linear -> split_bf16) belong to the outer call
            depth[0] += 1
            try:
                out = orig(*a, **k)
            finally:
                depth[0] -= 1
            if top:
                fam_name = family
                if name == 'linear':      # split the GEMM engine by regime: decoder M=100 chains are latency-bound
                    fam_name = 'gemm_m100' if int(np.prod(a[0].shape[:-1])) <= 128 else 'gemm_tokens'
                elif name == 'conv2d_nhwc':
                    fam_name = 'gemm_conv'
                elif name == 'mask_logits':
                    fam_name = 'gemm_mask_logits'
                calls.append((fam_name, orig, a, k, flops_fn(a, k) if flops_fn else 0.0,
                              bytes_fn(a, k) if bytes_fn else 0.0))
            return out
        setattr(ops, name, f)

    for s in spec:
        wrap(*s)
    try:
        runners = getattr(det, '_runners', None)
        det._runners = None
        xb = img[None].expand(batch, -1, -1, -1).contiguous()
        cls, mlr, _ = det.panoptic_head.simple_test_with_query(det.extract_feat(xb), [[meta]] * batch, upsample=False)
        from openpvsg_b200 import engine as _engine
        _engine.postprocess_batch(det, cls, mlr[:, 0].contiguous(), (736, 1280), (H, W), (H, W))
        torch.cuda.synchronize()
    finally:
        det._runners = runners
        for n, o in saved.items():
            setattr(ops, n, o)
    fam = {}
    for family in sorted(set(c[0] for c in calls)):
        mine = [c for c in calls if c[0] == family]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _, fn, a, k, _, _ in mine:
                fn(*a, **k)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _, fn, a, k, _, _ in mine:
                fn(*a, **k)
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        fam[family] = dict(ms=float(np.median(ts)) / batch, gflop=sum(c[4] for c in mine) / 1e9 / batch,
                           gbyte=sum(c[5] for c in mine) / 1e9 / batch, launches=len(mine) / batch)
        del g
    return fam


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--frames', type=int, default=100, help='frames per GPU per step')
    ap.add_argument('--distinct', type=int, default=100, help='distinct synthetic frames per rank (SURVEY 8d: 100 frames, seeds 0..99)')
    ap.add_argument('--cpu-frames', type=int, default=2, help='frames of the cpu_baseline sample')
    ap.add_argument('--batch', type=int, default=20, help='frames pushed through the network together (one graph replay); '
                    '100 frames per step = 5 replays of 20, measured 8 -> 322, 10 -> 342, 20 -> 348, 25 -> 347 frames/s')
    ap.add_argument('--backbone', default='r50', choices=['r50', 'swin_b'],
                    help='r50 = BASELINE configs[1] (the headline); swin_b = the Swin-B backbone BASELINE configs[2] names')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='eager launches (for ncu launch lists)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=240))   # fail fast on a lost rank

    import openpvsg_b200 as pv
    from openpvsg_b200 import configs, engine, lib, synthetic as syn, tubes
    peaks = load_peaks()
    swin = args.backbone == 'swin_b'
    det = pv.build_detector(configs.mask2former_swin(True) if swin else configs.mask2former_r50(True))
    det.load_state_dict(syn.mask2former_state_dict(seed=0, **SWIN_SD) if swin else syn.mask2former_state_dict(seed=0))
    det.to(dev)
    if not args.no_graph:
        engine.enable_cuda_graph(det)
    else:
        det._runners = None
    meta = syn.frame_meta(H, W)
    host = [f.pin_memory() for f in make_frames(args.distinct, rank)]
    resident = [f.to(dev) for f in host]
    host_batches = [torch.stack([host[(k + j) % len(host)] for j in range(args.batch)]).pin_memory()
                    for k in range(0, len(host), max(1, args.batch))] or [torch.stack(host).pin_memory()]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_step(frames, api):
        """One clip shard: every frame through the detector, then tube linking."""
        entries = []
        local_linker = tubes.TubeLinker() if world == 1 else None

        def consume(res):
            ids = list(res['query_feats'].keys())
            if local_linker is not None:
                # single process: link while the GPU works on the next batch (concat_seq is incremental)
                local_linker.add_frame(ids, [np.asarray(torch.as_tensor(res['query_feats'][k][0]).cpu()) for k in ids])
            else:
                # compact copy of the kept entries (the pinned ring is reused); they are exchanged once per clip
                entries.append((ids, np.stack([np.asarray(res['query_feats'][k][0]) for k in ids]) if ids
                                else np.zeros((0, 256), np.float32)))

        if api in ('sync', 'sync1') or det._runners is None:
            # the reference's call (public API; pinned host input when `api`): samples_per_gpu = batch
            # frames per call, as mmdet's single_gpu_test would feed them ('sync1': samples_per_gpu = 1)
            nb = args.batch if api == 'sync' else 1
            for i in range(0, args.frames, nb):
                xs = [frames[(i + j) % len(frames)] for j in range(min(nb, args.frames - i))]
                if api in ('sync', 'sync1'):
                    # the collated, pinned batch tensor a DataLoader(pin_memory=True) hands over
                    xb = host_batches[(i // nb) % len(host_batches)] if (len(xs) == nb and nb > 1) else torch.stack(xs)
                    xd = xb.to(dev, non_blocking=True)
                    out = det(return_loss=False, rescale=True, img=[xd], img_metas=[[dict(meta)] * len(xs)],
                              ref_img=[xd[:, None]], ref_img_metas=[[dict(meta)] for _ in xs])
                    for r in out:
                        consume(r[0])
                else:
                    res = det.simple_test(None, None, ref_img=xs[0][None, None], ref_img_metas=[[meta]],
                                          rescale=True)[0][0]
                    consume(res)
        elif api in ('pipelined', 'device'):
            # the public streaming call (pinned host frames: 'pipelined'; frames resident in HBM: 'device'): software-pipelined batches, half-sized first / last batch so that the host->device
            # fill and the device->host drain of a clip cost half a batch each (engine.batch_schedule)
            # (the graphs of the two runners are ordered by an event: left to run concurrently on their own streams they contend
            #  for the SMs -- tools/e2e_ab.py at 8 GPUs: ramp 274 ms, plain 287 ms, ramp with unordered runners 313 ms per step)
            engine.stream_frames(det, meta, [frames[i % len(frames)] for i in range(args.frames)], args.batch, consume)
        else:
            # same kernels, software-pipelined: frame i+1 is submitted before frame i is collected
            runner = engine.get_runner(det, meta, True, batch=args.batch)
            pend = None
            for i in range(0, args.frames, args.batch):
                nxt = runner.submit([frames[(i + j) % len(frames)] for j in range(min(args.batch, args.frames - i))])
                if pend is not None:
                    for r in runner.collect(pend, copy=False):
                        consume(r)
                pend = nxt
            for r in runner.collect(pend, copy=False):
                consume(r)
        if local_linker is not None:
            return local_linker
        return tubes.gather_and_link(entries, args.frames * world, device=dev)

    def timed(fn, steps, warmup):
        """`steps` calls of fn() between two barriers, device-timed (CUDA events on the current stream), max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        n0 = lib.launch_count[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, lib.launch_count[0] - n0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, linker, _ = timed(lambda: run_step(resident, 'device'), args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _, _ = timed(lambda: run_step(host, 'pipelined'), args.steps, 1)     # FrameRunner.submit/collect, pinned host frames
    sync_steps = max(1, args.steps // 2)
    ms_sync, _, _ = timed(lambda: run_step(host, 'sync'), sync_steps, 1)         # model(return_loss=False, ...), `batch` samples per call
    lat1 = None
    if det._runners is not None:
        # the reference's own test configuration: samples_per_gpu = 1, one synchronous call per frame
        fr = args.frames
        args.frames = min(20, fr)
        ms_one, _, _ = timed(lambda: run_step(host, 'sync1'), 1, 1)
        lat1 = ms_one / args.frames
        args.frames = fr
    if det._runners:
        per_frame = max(r.launches_per_frame for r in det._runners.values() if r.batch == args.batch)
    else:
        n0 = lib.launch_count[0]
        run_step(resident[:1] * 1, False) if args.frames == 1 else None
        per_frame = (lib.launch_count[0] - n0) or 638
    launches = int(per_frame * args.frames * args.steps) * world   # every rank launches the same work

    # ---- BASELINE configs[4] at this N: 380-frame clip (76 s @ 5 FPS) sharded over the ranks, VPS -> all-gather at
    # tube linking -> relation head on every rank; configs[2]: 300-frame Swin-B clip sharded the same way
    extra = {}
    if det._runners is not None and not swin:
        try:
            extra['end2end_clip'] = end2end_bench(det, meta, host, args.batch, dev, world, rank, timed,
                                                  with_cpu=not args.no_cpu_baseline)
        except Exception as ex:   # the headline metric must not depend on the auxiliary measurements
            extra['end2end_clip'] = dict(error=repr(ex))
        try:
            extra['swin_b_clip'] = swin_clip_bench(pv, configs, engine, syn, tubes, meta, host, args.batch, dev, world, rank, timed)
        except Exception as ex:
            extra['swin_b_clip'] = dict(error=repr(ex))

    # ---- SURVEY 8f rank 4 at this N: data-parallel training step (every rank takes part: gradient all-reduce)
    if not swin:
        try:
            extra['training_step'] = training_bench(dev, not args.no_cpu_baseline and world == 1, world=world, rank=rank)
        except Exception as ex:
            if world > 1:
                raise                      # a rank that drops out of a collective would hang the others
            extra['training_step'] = dict(error=repr(ex))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_frames = args.frames * world * args.steps
    value = total_frames / (ms_dev * 1e-3)
    e2e = total_frames / (ms_e2e * 1e-3)
    e2e_sync = args.frames * world * sync_steps / (ms_sync * 1e-3)
    fam = kernel_breakdown(det, resident[0], meta, args.batch)
    tot_ms = sum(d['ms'] for d in fam.values())
    gk = [k for k in fam if k.startswith('gemm')]
    g = dict(ms=sum(fam[k]['ms'] for k in gk) or 1.0, gflop=sum(fam[k]['gflop'] for k in gk),
             gbyte=sum(fam[k]['gbyte'] for k in gk), launches=sum(fam[k]['launches'] for k in gk))
    achieved_tf = g['gflop'] / g['ms']  # GFLOP / ms == TFLOP/s
    traffic, traffic_src = kernel_traffic('gemm')
    roofline = dict(kernel='gemm_tc_kernel (tcgen05 split-bf16 GEMM / implicit-GEMM conv engine; algorithmic 2MNK '
                           'flops, each product = 3 bf16 MMAs => ceiling peak/3 on algorithmic flops)', bound='tensor',
                    achieved=round(achieved_tf, 2), peak=peaks['tf_sustained'], unit='TFLOP/s',
                    frac=round(achieved_tf / peaks['tf_sustained'], 4), traffic=traffic, traffic_source=traffic_src,
                    frac_of_split_bf16_ceiling=round(3 * achieved_tf / peaks['tf_sustained'], 4),
                    peak_source=peaks['source'] + ', sustained bf16 figure (kernel timed inside a long step)',
                    launches_per_frame=round(g['launches'], 1), gflop_per_frame=round(g['gflop'], 1),
                    gflop_per_launch=round(g['gflop'] / max(g['launches'], 1e-9), 2),
                    us_per_launch=round(1e3 * g['ms'] / max(g['launches'], 1e-9), 2),
                    algorithmic_bytes_per_frame=int(g['gbyte'] * 1e9),
                    algorithmic_bytes_per_launch=int(g['gbyte'] * 1e9 / max(g['launches'], 1e-9)),
                    algorithmic_bytes_note='fp32-equivalent operand + result bytes of every call of the family (activations, '
                                           'weights, residual, output; 4 B per element: a (hi, lo) bf16 plane pair is the same size)',
                    achieved_GBps_algorithmic=round(g['gbyte'] / (g['ms'] * 1e-3), 1),
                    share_of_frame=round(g['ms'] / tot_ms, 3))
    m = fam.get('msda')
    kernels = {k: dict(ms_per_frame=round(d['ms'], 3), share=round(d['ms'] / tot_ms, 3), launches=round(d['launches'], 1))
               for k, d in fam.items()}
    for k in gk:
        if fam[k]['gflop'] > 0:
            kernels[k]['TFLOPs'] = round(fam[k]['gflop'] / fam[k]['ms'], 1)
    if m:
        gbps = m['gbyte'] / (m['ms'] * 1e-3)
        kernels['msda']['achieved_GBps'] = round(gbps, 1)
        kernels['msda']['frac_hbm'] = round(gbps / peaks['hbm_gbs'], 4)
        mt, mt_src = kernel_traffic('msda')
        # north_star target: MSDeformAttn as a fraction of the HBM roofline (algorithmic bytes: value + raw
        # projections + output = 3200 B per token and layer)
        extra['roofline_msda'] = dict(kernel=MSDA_KERNEL, bound='hbm', achieved=round(gbps, 1),
                                      peak=peaks['hbm_gbs'], unit='GB/s', frac=round(gbps / peaks['hbm_gbs'], 4),
                                      traffic=mt, traffic_source=mt_src,
                                      algorithmic_bytes_per_launch=int(m['gbyte'] * 1e9 * args.batch / max(m['launches'] * args.batch, 1e-9)),
                                      us_per_launch=round(1e3 * m['ms'] / max(m['launches'], 1e-9), 2),
                                      note=MSDA_NOTE)
    ml = fam.get('gemm_mask_logits')
    if ml and ml['gflop'] > 0:
        tf = ml['gflop'] / ml['ms']
        gbps = ml['gbyte'] / (ml['ms'] * 1e-3)
        extra['roofline_mask_einsum'] = dict(
            kernel='gemm_tc_kernel (pvsg_mask_logits)', bound='hbm', unit='GB/s', achieved=round(gbps, 1),
            peak=peaks['hbm_gbs'], frac=round(gbps / peaks['hbm_gbs'], 4),
            algorithmic_bytes_per_frame=int(ml['gbyte'] * 1e9), tflops=round(tf, 1),
            tensor_frac=round(tf / peaks['tf_sustained'], 4),
            note='AI = 36 FLOP/B (fp32 I/O, Q = 100 rows) << ridge 250: HBM-bound by construction (SURVEY.md 8d); bytes = embed + '
                 'feature operands + the result actually written (fp32 logits for the last layer, 1-byte sign masks for the nine '
                 'intermediate calls, which run on pooled features: exact, 3x fewer flops)')
    cpu = None
    if not args.no_cpu_baseline:
        fps, times = cpu_oracle_fps(args.cpu_frames, swin=swin)
        cpu = dict(value=fps, unit='frames/s', cores=os.cpu_count(), kind='port',
                   sample=f'{args.cpu_frames} frames @720p through the CPU oracle (torch CPU fp32, '
                          f'{os.cpu_count()} threads), 1 warm-up frame')
    try:
        extra['relation_head'] = relation_bench(dev, not args.no_cpu_baseline)
    except Exception as ex:
        extra['relation_head'] = dict(error=repr(ex))
    try:
        extra['relation_set'] = relset_bench(dev, not args.no_cpu_baseline)
        extra['relation_set']['frac_hbm'] = round(extra['relation_set']['achieved_GBps'] / peaks['hbm_gbs'], 4)
    except Exception as ex:
        extra['relation_set'] = dict(error=repr(ex))
    try:
        if det._runners is not None:
            extra['tube_dump'] = tube_dump_bench(det, meta, resident, args.batch)
    except Exception as ex:
        extra['tube_dump'] = dict(error=repr(ex))
    in_bytes = 3 * 736 * 1280 * 4
    out_bytes = H * W * 4 + (1 + 400) * 4 + 10 * H * W + 100 * 256 * 4
    line = dict(metric=METRIC.replace('R50', 'Swin-B') if swin else METRIC, value=round(value, 3), unit='frames/s', n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=round(ms_dev / args.steps, 3), higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype=DTYPE, data='synthetic',
                config=workload_config(args, world), tubes=len(linker.object_list),
                e2e=dict(value=round(e2e, 3), unit='frames/s', h2d_bytes_per_step=in_bytes * args.frames * world,
                         d2h_bytes_per_step=out_bytes * args.frames * world, ms_per_step=round(ms_e2e / args.steps, 3),
                         api='engine.stream_frames (FrameRunner.submit/collect pipelined, half-sized first / last batch) on pinned host frames',
                         sync_api_value=round(e2e_sync, 3),
                         sync_api=f'model(return_loss=False, rescale=True, img=..., ref_img=...) with {args.batch} samples per call, synchronous',
                         latency_ms_batch1=None if lat1 is None else round(lat1, 3),
                         latency_api='model(return_loss=False, rescale=True, ...) with samples_per_gpu = 1 (the reference test '
                                     'config), pinned host frame in, result dict out, synchronous; mean over 20 frames'),
                gpu_launches=launches, clocks=clocks, roofline=roofline, kernels=kernels, cpu_baseline=cpu, **extra)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
