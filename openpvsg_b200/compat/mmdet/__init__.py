"""Minimal stand-in for the parts of mmdet 2.25 that tools/test.py touches (see ../README.md)."""
__version__ = '2.25.0+openpvsg_b200.compat'
