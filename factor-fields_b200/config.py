"""Config layer: the reference merges `configs/defaults.yaml` <- `<yaml>` <- dot-list CLI overrides with
OmegaConf (train_per_scene.py:243-247) and hands an attribute-style object to the model, which reads it
lazily and *writes* `cfg.dataset.aabb` (FactorFields.py:808).  OmegaConf is not available here, so this is a
small equivalent: PyYAML with a float resolver for `1e-3`-style scalars (OmegaConf parses them as floats,
stock PyYAML as strings) and an attribute dict."""
import ast
import os
import re

import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'configs')


class _Loader(yaml.SafeLoader):
    pass


_Loader.add_implicit_resolver(
    'tag:yaml.org,2002:float',
    re.compile(r'''^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                    |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                    |\.[0-9_]+(?:[eE][-+][0-9]+)?
                    |[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$''', re.X),
    list('-+0123456789.'))


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        import copy
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _wrap(d):
    if isinstance(d, dict):
        return AttrDict({k: _wrap(v) for k, v in d.items()})
    return d


def merge_cfg(a, b):
    out = AttrDict(a)
    for k, v in b.items():
        if k in out and isinstance(out[k], dict) and isinstance(v, dict):
            out[k] = merge_cfg(out[k], v)
        else:
            out[k] = _wrap(v)
    return out


def _read(path):
    with open(path) as f:
        return _wrap(yaml.load(f, Loader=_Loader) or {})


def apply_dotlist(cfg, dotlist):
    """`model.basis_type=vm training.n_iters=100` style overrides."""
    for item in dotlist:
        key, _, val = item.partition('=')
        try:       # OmegaConf.from_cli gives dot-list values YAML semantics: true / false / null / 1e-3 / [1, 2] / bare strings
            val = yaml.load(val, Loader=_Loader)
        except Exception:
            try:
                val = ast.literal_eval(val)
            except Exception:
                pass
        node = cfg
        parts = key.split('.')
        for p in parts[:-1]:
            if p not in node:
                node[p] = AttrDict()
            node = node[p]
        node[parts[-1]] = _wrap(val)
    return cfg


def load_cfg(name_or_path, dotlist=(), config_dir=None):
    """defaults.yaml <- <yaml> <- dotlist.  `name_or_path` may be a file in configs/ or a path."""
    cdir = config_dir or CONFIG_DIR
    path = name_or_path if os.path.exists(name_or_path) else os.path.join(cdir, name_or_path)
    base = _read(os.path.join(cdir, 'defaults.yaml'))
    cfg = merge_cfg(base, _read(path))
    return apply_dotlist(cfg, dotlist)
