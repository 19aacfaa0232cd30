"""mmdet.apis.single_gpu_test / multi_gpu_test (mmdet/apis/test.py): iterate the loader, call
model(return_loss=False, rescale=True, **data), extend the result list; and the training entry points of
mmdet/apis/train.py (init_random_seed, set_random_seed, train_detector) for tools/train.py."""
import os
import random
import time

import numpy as np
import torch


def single_gpu_test(model, data_loader, show=False, out_dir=None, show_score_thr=0.3):
    if show or out_dir:
        raise NotImplementedError('compat single_gpu_test: visualisation is outside the B200 backend')
    model.eval()
    results = []
    for data in data_loader:
        with torch.no_grad():
            result = model(return_loss=False, rescale=True, **data)
        results.extend(result)
    return results


def multi_gpu_test(model, data_loader, tmpdir=None, gpu_collect=False):
    """One rank per GPU; every rank runs its shard of the loader, results are gathered on rank 0 in dataset order."""
    import torch.distributed as dist
    results = single_gpu_test(model, data_loader)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return results
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, results)
    if dist.get_rank() != 0:
        return None
    ordered = []
    for i in range(max(len(p) for p in parts)):
        for p in parts:
            if i < len(p):
                ordered.append(p[i])
    return ordered[:len(data_loader.dataset)]


# ---------------------------------------------------------------------------------------------------- training
def init_random_seed(seed=None, device='cuda'):
    """mmdet.apis.init_random_seed: a given seed is returned as is; otherwise rank 0 draws one and broadcasts it."""
    import torch.distributed as dist
    if seed is not None:
        return int(seed)
    seed = np.random.randint(2 ** 31)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor(seed if dist.get_rank() == 0 else 0, dtype=torch.int32, device=device)
        dist.broadcast(t, src=0)
        seed = int(t.item())
    return int(seed)


def set_random_seed(seed, deterministic=False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


_NORMS = (torch.nn.LayerNorm, torch.nn.GroupNorm, torch.nn.modules.batchnorm._BatchNorm)


def build_optimizer(model, cfg):
    """mmcv DefaultOptimizerConstructor for the options of _base_/schedules/m2f_schedules.py: ``paramwise_cfg.custom_keys``
    (the longest key that is a substring of the parameter name wins: lr_mult / decay_mult) and ``norm_decay_mult`` for the
    parameters of normalisation layers; parameters with requires_grad=False are left out."""
    cfg = dict(cfg)
    typ = cfg.pop('type')
    pw = dict(cfg.pop('paramwise_cfg', None) or {})
    custom = dict(pw.get('custom_keys', {}))
    keys = sorted(custom, key=lambda k: (-len(k), k))
    norm_decay = pw.get('norm_decay_mult')
    base_lr, base_wd = cfg['lr'], cfg.get('weight_decay', 0.0)
    norm_params = {id(p) for m in model.modules() if isinstance(m, _NORMS) for p in m.parameters(recurse=False)}
    groups = []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        g = dict(params=[p], name=name, lr=base_lr, weight_decay=base_wd)
        hit = next((k for k in keys if k in name), None)
        if hit is not None:
            g['lr'] = base_lr * custom[hit].get('lr_mult', 1.0)
            g['weight_decay'] = base_wd * custom[hit].get('decay_mult', 1.0)
        elif norm_decay is not None and id(p) in norm_params:
            g['weight_decay'] = base_wd * norm_decay
        g['initial_lr'] = g['lr']
        groups.append(g)
    return getattr(torch.optim, typ)(groups, **cfg)


def _lr_factor(lr_cfg, epoch, it):
    """mmcv StepLrUpdaterHook (+ linear warm-up by iteration): the multiplier on every group's initial lr."""
    lr_cfg = lr_cfg or {}
    factor = 1.0
    if lr_cfg.get('policy', 'fixed') == 'step':
        steps = lr_cfg['step'] if isinstance(lr_cfg['step'], (list, tuple)) else [lr_cfg['step']]
        factor = lr_cfg.get('gamma', 0.1) ** sum(epoch >= s for s in steps)
    if lr_cfg.get('warmup') == 'linear' and it < lr_cfg.get('warmup_iters', 0):
        k = (1 - it / lr_cfg['warmup_iters']) * (1 - lr_cfg.get('warmup_ratio', 0.1))
        factor *= 1 - k
    return factor


def train_detector(model, dataset, cfg, distributed=False, validate=False, timestamp=None, meta=None):
    """mmdet.apis.train_detector, single process: data loader -> optimizer (build_optimizer) -> EpochBasedRunner /
    IterBasedRunner loop of ``model.train_step`` -> backward -> gradient clipping (``optimizer_config.grad_clip``) ->
    optimizer step, with the step lr policy + linear warm-up, text logging every ``log_config.interval`` iterations and a
    checkpoint per ``checkpoint_config.interval`` epochs in ``cfg.work_dir``.  ``distributed=True``: one process per GPU (the
    launcher has initialised torch.distributed), a DistributedSampler shards the clips, the weights are broadcast once and the
    gradients averaged by ``openpvsg_b200.dist_train.allreduce_gradients`` after every backward.  Evaluation hooks
    (``validate``) and wandb are outside this floor."""
    from mmdet.datasets import build_dataloader
    from mmdet.utils import get_root_logger
    from openpvsg_b200 import dist_train
    logger = get_root_logger(log_level=cfg.get('log_level', 'INFO'))
    if validate:
        logger.warning('compat train_detector: evaluation during training is not part of this floor; continuing without it')
    dataset = dataset[0] if isinstance(dataset, (list, tuple)) else dataset
    data_cfg = cfg.get('data', {})
    loader = build_dataloader(dataset, samples_per_gpu=data_cfg.get('samples_per_gpu', 1),
                              workers_per_gpu=data_cfg.get('workers_per_gpu', 0), shuffle=True, seed=cfg.get('seed'), train=True,
                              dist=distributed)
    device = cfg.get('device', 'cuda')
    model.to(device)
    model.train()
    if distributed:                          # one process per GPU: same start, clips sharded by the sampler, gradients averaged
        dist_train.broadcast_parameters(model)
    optimizer = build_optimizer(model, cfg.optimizer)
    bucket = None
    clip = (cfg.get('optimizer_config') or {}).get('grad_clip')
    runner_cfg = cfg.get('runner') or dict(type='EpochBasedRunner', max_epochs=cfg.get('total_epochs', 1))
    by_epoch = runner_cfg.get('type', 'EpochBasedRunner') == 'EpochBasedRunner'
    max_epochs = runner_cfg.get('max_epochs', 1) if by_epoch else 10 ** 9
    max_iters = runner_cfg.get('max_iters', 10 ** 12) if not by_epoch else 10 ** 12
    interval = (cfg.get('log_config') or {}).get('interval', 50)
    ckpt = cfg.get('checkpoint_config')
    params = [p for g in optimizer.param_groups for p in g['params']]
    it, history = 0, []
    t0 = time.time()
    for epoch in range(max_epochs):
        if distributed and hasattr(loader.sampler, 'set_epoch'):
            loader.sampler.set_epoch(epoch)
        for data in loader:
            factor = _lr_factor(cfg.get('lr_config'), epoch, it)
            for g in optimizer.param_groups:
                g['lr'] = g['initial_lr'] * factor
            data = {k: _to_device(v, device) for k, v in data.items()}
            optimizer.zero_grad(set_to_none=True)
            out = model.train_step(data, optimizer)
            out['loss'].backward()
            if distributed:
                bucket = dist_train.allreduce_gradients(params, bucket)
            grad_norm = None
            if clip:
                grad_norm = float(torch.nn.utils.clip_grad_norm_(params, clip['max_norm'], clip.get('norm_type', 2)))
            optimizer.step()
            it += 1
            history.append(out['log_vars']['loss'])
            if it % interval == 0 or it == 1:
                logger.info(f'Epoch [{epoch + 1}][{it}]\tlr: {optimizer.param_groups[0]["lr"]:.3e}, loss: {history[-1]:.4f}, '
                            f'grad_norm: {grad_norm}, time: {(time.time() - t0) / it:.3f} s/iter')
            if it >= max_iters:
                break
        if ckpt is not None and by_epoch and (epoch + 1) % ckpt.get('interval', 1) == 0 and cfg.get('work_dir') and \
                (not distributed or dist_train.dist.get_rank() == 0):
            os.makedirs(cfg.work_dir, exist_ok=True)
            torch.save(dict(meta=dict(ckpt.get('meta', {}) or {}, epoch=epoch + 1, iter=it, **(meta or {})),
                            state_dict=model.state_dict(), optimizer=optimizer.state_dict()),
                       os.path.join(cfg.work_dir, f'epoch_{epoch + 1}.pth'))
        if it >= max_iters:
            break
    return dict(iters=it, loss_history=history, optimizer=optimizer)


def _to_device(v, device):
    if torch.is_tensor(v):
        return v.to(device, non_blocking=True)
    if isinstance(v, (list, tuple)):
        return type(v)(_to_device(x, device) for x in v)
    return v
