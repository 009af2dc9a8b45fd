"""Minimal stand-in for the reference's `dnnlib` package: only what the hot path touches
(`EasyDict`, `util.construct_class_by_name`; reference dnnlib/util.py:41,303)."""
from . import util
from .util import EasyDict

__all__ = ['util', 'EasyDict']
