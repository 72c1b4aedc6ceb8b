"""factor-fields_b200 — B200-native query-and-render hot path of Factor Fields.

Import as `ffb200` (the repo-root shim `ffb200.py` loads this directory under that name, because the
directory name is not a valid Python identifier)."""
from . import native  # noqa: F401
from .config import load_cfg, merge_cfg, AttrDict  # noqa: F401

__all__ = ['native', 'load_cfg', 'merge_cfg', 'AttrDict']
