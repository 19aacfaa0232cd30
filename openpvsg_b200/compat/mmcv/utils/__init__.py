"""mmcv.utils names tools/train.py imports."""


def get_git_hash(fallback='unknown', digits=None):
    return fallback if digits is None else fallback[:digits]
