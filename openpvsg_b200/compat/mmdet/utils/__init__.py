"""mmdet.utils helpers tools/test.py calls before building the model."""
import torch
import torch.nn as nn


def replace_cfg_vals(cfg):
    return cfg


def update_data_root(cfg, logger=None):
    return None


def compat_cfg(cfg):
    """mmdet.utils.compat_cfg: make sure data.test_dataloader exists (old configs keep the loader keys in `data`)."""
    data = cfg.get('data')
    if data is not None and 'test_dataloader' not in data:
        data['test_dataloader'] = type(data)()
    return cfg


def setup_multi_processes(cfg):
    return None


def get_device():
    return 'cuda' if torch.cuda.is_available() else 'cpu'


class _SingleDevice(nn.Module):
    """What MMDataParallel does for a single GPU: move the batch to the device, call the module."""

    def __init__(self, module, device):
        super().__init__()
        self.module = module.to(device)
        self.device = torch.device(device)

    def _move(self, x):
        if torch.is_tensor(x):
            return x.to(self.device, non_blocking=True)
        if isinstance(x, list):
            return [self._move(v) for v in x]
        if isinstance(x, tuple):
            return tuple(self._move(v) for v in x)
        return x

    def forward(self, *args, **kwargs):
        return self.module(*[self._move(a) for a in args], **{k: self._move(v) for k, v in kwargs.items()})


def build_dp(model, device='cuda', dim=0, device_ids=None, **kwargs):
    index = (device_ids or [0])[0]
    return _SingleDevice(model, f'cuda:{index}' if device == 'cuda' else device)


def build_ddp(model, device='cuda', device_ids=None, **kwargs):
    """Inference needs no gradient synchronisation: one replica per rank on its own GPU."""
    index = (device_ids or [0])[0]
    return _SingleDevice(model, f'cuda:{index}' if device == 'cuda' else device)


def collect_env():
    import sys
    return {'sys.platform': sys.platform, 'Python': sys.version.replace('\n', ''), 'PyTorch': torch.__version__,
            'CUDA available': torch.cuda.is_available(),
            'GPU 0': torch.cuda.get_device_name(0) if torch.cuda.is_available() else None, 'backend': 'openpvsg_b200'}


def get_root_logger(log_file=None, log_level='INFO'):
    import logging
    logger = logging.getLogger('mmdet')
    if not logger.handlers:
        logger.addHandler(logging.StreamHandler())
    if log_file is not None and not any(getattr(h, 'baseFilename', None) == log_file for h in logger.handlers):
        logger.addHandler(logging.FileHandler(log_file, 'w'))
    for h in logger.handlers:
        h.setFormatter(logging.Formatter('%(asctime)s - %(name)s - %(levelname)s - %(message)s'))
    logger.setLevel(getattr(logging, log_level) if isinstance(log_level, str) else log_level)
    return logger
