"""mmcv-style registries and a minimal config loader.

The reference's hot path sits behind ``mmcv.utils.Registry`` + ``build_from_cfg(type=...)``
(e.g. models/mask2former_vps/mask2former.py:33,
mask2former_video_head.py:20, position_encoding.py:9).  mmcv / mmdet are not installable
here, so the same mechanism is restated in ~100 lines: ``type`` strings of the
reference's configs resolve to the B200 modules registered under the reference's names.
INTEGRATION.md shows how the same classes register into the real mmdet registries.
"""
import copy
import importlib.util
import os


class ConfigDict(dict):
    """dict with attribute access (mmcv ``ConfigDict`` behaviour used by the reference:
    ``transformer_decoder.transformerlayers.attn_cfgs.num_heads``,
    mask2former_video_head.py:86-87)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_cfg(obj):
    if isinstance(obj, dict):
        return ConfigDict({k: to_cfg(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_cfg(v) for v in obj)
    return obj


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f'{key} is already registered in {self.name}')
            self._modules[key] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self._modules.get(key)

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)

    def __contains__(self, key):
        return key in self._modules


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict) or 'type' not in cfg:
        raise KeyError(f'cfg must be a dict with a "type" key, got {cfg}')
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    typ = args.pop('type')
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f'{typ} is not in the {registry.name} registry')
    return cls(**args)


BACKBONES = Registry('backbone')
HEADS = Registry('head')
DETECTORS = Registry('detector')
PLUGIN_LAYERS = Registry('plugin layer')
ATTENTION = Registry('attention')
FEEDFORWARD_NETWORK = Registry('feed-forward network')
TRANSFORMER_LAYER = Registry('transformer layer')
TRANSFORMER_LAYER_SEQUENCE = Registry('transformer layer sequence')
POSITIONAL_ENCODING = Registry('position encoding')


def build_backbone(cfg):
    return BACKBONES.build(to_cfg(cfg))


def build_head(cfg):
    return HEADS.build(to_cfg(cfg))


def build_detector(cfg, train_cfg=None, test_cfg=None):
    """mmdet.models.build_detector (tools/test.py:226, prepare_query_tube_vps.py:214)."""
    cfg = to_cfg(cfg)
    extra = {}
    if train_cfg is not None:
        extra['train_cfg'] = train_cfg
    if test_cfg is not None:
        extra['test_cfg'] = test_cfg
    return DETECTORS.build(cfg, extra)


def build_plugin_layer(cfg):
    cfg = to_cfg(cfg)
    layer = PLUGIN_LAYERS.build(cfg)
    return cfg['type'], layer


def build_transformer_layer_sequence(cfg):
    return TRANSFORMER_LAYER_SEQUENCE.build(to_cfg(cfg))


def build_positional_encoding(cfg):
    return POSITIONAL_ENCODING.build(to_cfg(cfg))


# --------------------------------------------------------------------------------------
# mmcv.Config.fromfile subset: python config files with `_base_` inheritance
# --------------------------------------------------------------------------------------
def _merge(base, child):
    out = copy.deepcopy(base)
    for k, v in child.items():
        if isinstance(v, dict) and v.pop('_delete_', False):
            out[k] = v
        elif isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def load_config(path):
    """Evaluate an mmcv-style python config (``configs/mask2former_vps/*.py``) including its
    ``_base_`` list, child keys overriding base keys recursively."""
    path = os.path.abspath(path)
    spec = importlib.util.spec_from_file_location('_pvsg_cfg', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = {k: v for k, v in vars(mod).items() if not k.startswith('__') and
           not isinstance(v, type(os))}
    bases = cfg.pop('_base_', [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        merged = _merge(merged, load_config(os.path.join(os.path.dirname(path), b)))
    return to_cfg(_merge(merged, cfg))
