"""Minimal stand-in for the parts of mmcv 1.4 that tools/test.py touches (see ../README.md)."""
import argparse
import json
import os
import pickle

from openpvsg_b200.registry import ConfigDict, load_config, to_cfg

__version__ = '1.4.0+openpvsg_b200.compat'


class Config(ConfigDict):
    """mmcv.Config subset: python config files with `_base_` inheritance, attribute access, merge_from_dict."""

    def __init__(self, cfg_dict=None, **kwargs):
        super().__init__(to_cfg(dict(cfg_dict or {}, **kwargs)))      # nested dicts get attribute access too

    @staticmethod
    def fromfile(filename, **kwargs):
        cfg = Config(load_config(filename))
        cfg['filename'] = filename
        return cfg

    @property
    def pretty_text(self):
        import pprint
        return pprint.pformat(_plain(self), width=120)

    def dump(self, file=None):
        """mmcv.Config.dump: the merged config as a python file of top-level assignments."""
        text = ''.join(f'{k} = {_plain(v)!r}\n' for k, v in self.items() if k != 'filename')
        if file is None:
            return text
        with open(file, 'w') as f:
            f.write(text)

    def merge_from_dict(self, options):
        for dotted, value in options.items():
            node = self
            keys = dotted.split('.')
            for k in keys[:-1]:
                if k not in node or not isinstance(node[k], dict):
                    node[k] = ConfigDict()
                node = node[k]
            node[keys[-1]] = to_cfg(value)


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return type(v)(_plain(x) for x in v) if not isinstance(v, range) else list(v)
    return v


class DictAction(argparse.Action):
    """`--cfg-options a.b=1 c=[1,2] d=true` -> {'a.b': 1, 'c': [1, 2], 'd': True}."""

    @staticmethod
    def _parse(val):
        for cast in (int, float):
            try:
                return cast(val)
            except ValueError:
                pass
        if val.lower() in ('true', 'false'):
            return val.lower() == 'true'
        if val == 'None':
            return None
        if (val.startswith('[') and val.endswith(']')) or (val.startswith('(') and val.endswith(')')):
            inner = [v for v in val[1:-1].split(',') if v != '']
            items = [DictAction._parse(v.strip()) for v in inner]
            return items if val[0] == '[' else tuple(items)
        return val.strip('\'"')

    def __call__(self, parser, namespace, values, option_string=None):
        options = {}
        for kv in values:
            key, val = kv.split('=', maxsplit=1)
            options[key] = self._parse(val)
        setattr(namespace, self.dest, options)


def mkdir_or_exist(dir_name, mode=0o777):
    if dir_name:
        os.makedirs(os.path.expanduser(dir_name), mode=mode, exist_ok=True)


def dump(obj, file, **kwargs):
    if str(file).endswith('.json'):
        with open(file, 'w') as f:
            json.dump(obj, f)
    else:
        with open(file, 'wb') as f:
            pickle.dump(obj, f)


def load(file, **kwargs):
    if str(file).endswith('.json'):
        with open(file) as f:
            return json.load(f)
    with open(file, 'rb') as f:
        return pickle.load(f)
