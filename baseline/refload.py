# Baseline-measurement helper written at survey time. NOT product source and NOT a port:
# it only imports the UNMODIFIED reference checkout (FF_REF env var, else
# baseline/_ref/factor-fields, else /root/reference) in a container that lacks
# omegaconf/skimage/plyfile/imageio/kornia, by stubbing those non-hot-path imports.
import os, sys, types, re
import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FF_REF') or next(
    (p for p in (os.path.join(_HERE, '_ref', 'factor-fields'), '/root/reference') if os.path.isdir(p)), None)
assert REF, 'reference checkout not found: copy /root/reference to baseline/_ref/factor-fields (gitignored) or set FF_REF'
sys.path.insert(0, REF)
for name in ['skimage', 'skimage.measure', 'skimage.morphology', 'plyfile', 'imageio', 'kornia', 'lpips']:
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules['skimage'].measure = sys.modules['skimage.measure']
sys.modules['skimage'].morphology = sys.modules['skimage.morphology']
sys.modules['kornia'].create_meshgrid = None

# OmegaConf parses "1e-3" as float; stock PyYAML does not -> patch the resolver.
_L = yaml.SafeLoader
_L.add_implicit_resolver(
    u'tag:yaml.org,2002:float',
    re.compile(u'''^(?:[-+]?(?:[0-9][0-9_]*)\\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                   |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                   |\\.[0-9_]+(?:[eE][-+][0-9]+)?
                   |[-+]?\\.(?:inf|Inf|INF)|\\.(?:nan|NaN|NAN))$''', re.X),
    list(u'-+0123456789.'))


class AD(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def wrap(d):
    if isinstance(d, dict):
        return AD({k: wrap(v) for k, v in d.items()})
    return d


def merge(a, b):
    out = AD(a)
    for k, v in b.items():
        if k in out and isinstance(out[k], dict) and isinstance(v, dict):
            out[k] = merge(out[k], v)
        else:
            out[k] = v
    return out


def load_cfg(name):
    base = wrap(yaml.load(open(os.path.join(REF, 'configs', 'defaults.yaml')), Loader=_L))
    sec = wrap(yaml.load(open(os.path.join(REF, 'configs', name)), Loader=_L))
    return merge(base, sec)
