"""mmcv.runner subset used by tools/test.py: dist info, checkpoint loading."""
import os

import torch


def get_dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_dist(launcher, backend='nccl', **kwargs):
    import torch.distributed as dist
    if launcher != 'pytorch':
        raise NotImplementedError(f'compat init_dist: launcher {launcher!r} (use "pytorch" = torchrun, one rank per GPU)')
    rank = int(os.environ['RANK'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank % max(torch.cuda.device_count(), 1))))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group(backend=backend, **kwargs)


def load_checkpoint(model, filename, map_location=None, strict=False, logger=None, revise_keys=None):
    """mmcv load_checkpoint: accepts a bare state_dict or {'state_dict': ..., 'meta': ...}; strips a 'module.' prefix."""
    ckpt = torch.load(filename, map_location=map_location or 'cpu', weights_only=False)
    sd = ckpt.get('state_dict', ckpt) if isinstance(ckpt, dict) else ckpt
    sd = {(k[7:] if k.startswith('module.') else k): v for k, v in sd.items()}
    res = model.load_state_dict(sd, strict=strict)
    if res.missing_keys or res.unexpected_keys:
        print(f'load_checkpoint: missing {list(res.missing_keys)[:5]}..., unexpected {list(res.unexpected_keys)[:5]}...')
    return ckpt if isinstance(ckpt, dict) and 'state_dict' in ckpt else dict(state_dict=sd, meta={})


def wrap_fp16_model(model):
    raise NotImplementedError('fp16 wrapping: the B200 backend computes in split-bf16 with fp32 accumulation; remove `fp16` from the config')
