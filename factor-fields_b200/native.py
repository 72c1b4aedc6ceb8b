"""ctypes binding of libffb200.so (the C ABI declared in include/ffb200.h).

There is deliberately no fallback: if the shared library is missing, or an entry point fails, a
RuntimeError is raised.  Tensors only provide device memory and the current stream; every pointer
handed to the library is validated here (device, dtype, contiguity).
"""
import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, 'libffb200.so')
CSRC = os.path.join(_HERE, 'csrc')
HEADER = os.path.join(_ROOT, 'include', 'ffb200.h')

MAX_OPS, MAX_TERMS, MAX_FREQ = 64, 40, 16
MAP_IDS = {'sawtooth': 0, 'triangle': 1, 'sinc': 2, 'trigonometric': 3, 'x': 4}

c_f32p = C.c_void_p
i32, i64, f32 = C.c_int32, C.c_int64, C.c_float


class GatherOp(C.Structure):
    _fields_ = [('data', C.c_void_p), ('grad', C.c_void_p), ('C', i32), ('nd', i32), ('size', i32 * 3), ('src', i32 * 3),
                ('cst', f32 * 3), ('space', i32), ('level', i32), ('align_corners', i32), ('border', i32), ('nearest', i32)]


class Term(C.Structure):
    _fields_ = [('n_ops', i32), ('op', i32 * 3), ('col', i32)]


class FieldDesc(C.Structure):
    _fields_ = [('xdim', i32), ('in_dim', i32), ('aabb_min', f32 * 3), ('aabb_max', f32 * 3), ('mapping', i32), ('n_freq', i32),
                ('freq', f32 * MAX_FREQ), ('n_ops', i32), ('ops', GatherOp * MAX_OPS), ('n_cterms', i32), ('n_bterms', i32),
                ('cterms', Term * MAX_TERMS), ('bterms', Term * MAX_TERMS), ('coeff_width', i32), ('basis_width', i32),
                ('basis_is_x', i32), ('basis_perm', C.c_void_p)]


class SamplerDesc(C.Structure):
    _fields_ = [('aabb_min', f32 * 3), ('aabb_max', f32 * 3), ('step_size', f32), ('n_samples', i32), ('alpha_volume', C.c_void_p),
                ('alpha_size', i32 * 3), ('alpha_aabb_min', f32 * 3), ('alpha_inv_size', f32 * 3), ('alpha_thres', f32),
                ('mode', i32), ('bg_len', f32), ('z_table', C.c_void_p), ('alpha_outside', i32)]


class CompositeDesc(C.Structure):
    _fields_ = [('density_shift', f32), ('softplus', i32), ('distance_scale', f32), ('weight_thres', f32), ('white_bg', i32), ('white_bg_dev', C.c_void_p)]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into libffb200.so (in-tree, so it travels to the GPU box).  One object per
    source, compiled in parallel and re-used while it is newer than the source and every header."""
    from concurrent.futures import ThreadPoolExecutor
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))] + [HEADER]
    hdr_time = max(os.path.getmtime(p) for p in hdrs)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objdir = os.path.join(_HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    flags = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC']
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append([nvcc] + flags + ['-c', src, '-o', obj])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(o) for o in objs):
        run([nvcc, '-shared', '-o', LIB_PATH] + objs + ['-lcuda'])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                               '(there is no CPU or PyTorch fallback for the ffb200 kernels)')
        _lib = C.CDLL(LIB_PATH)
        _lib.ffb_last_error.restype = C.c_char_p
        _lib.ffb_launch_count.restype = C.c_uint64
        _lib.ffb_rgbmlp_workspace_bytes.restype = C.c_int64
        _lib.ffb_rgbmlp_stream_bytes.restype = C.c_int64
        if _lib.ffb_abi_version() != 1:
            raise RuntimeError('libffb200.so ABI version mismatch')
        for kv in filter(None, os.environ.get('FFB_TUNING', '').split(',')):      # experiment knobs: FFB_TUNING=key=value,key=value
            k, v = kv.split('=')
            if _lib.ffb_set_tuning(k.strip().encode(), int(v)) != 0:
                raise RuntimeError(f'ffb200: unknown tuning key {k!r} in FFB_TUNING')
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f'ffb200: {lib().ffb_last_error().decode()} (code {rc})')


def launch_count():
    return int(lib().ffb_launch_count())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def on_device_of(argpos):
    """Decorator for the op entry points: run with the CUDA device of positional argument `argpos` (a tensor) current, so
    that `FactorFields(cfg, 'cuda:1')` works without a global torch.cuda.set_device(1) — the library launches on the
    current device and `stream()` is that device's current stream."""
    def deco(fn):
        import functools

        @functools.wraps(fn)
        def wrapper(*a, **k):
            t = a[argpos]
            if torch.is_tensor(t) and t.is_cuda and t.device.index != torch.cuda.current_device():
                with torch.cuda.device(t.device):
                    return fn(*a, **k)
            return fn(*a, **k)
        return wrapper
    return deco


def ptr(t, dtype=torch.float32, allow_none=False, contiguous=True):
    if t is None:
        if allow_none:
            return C.c_void_p(0)
        raise RuntimeError('ffb200: required tensor is None')
    if not t.is_cuda:
        raise RuntimeError('ffb200: tensor must live on a CUDA device (there is no CPU path)')
    if t.dtype != dtype:
        raise RuntimeError(f'ffb200: expected dtype {dtype}, got {t.dtype}')
    if contiguous and not t.is_contiguous():
        raise RuntimeError('ffb200: tensor must be contiguous')
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f'ffb200: tensor lives on {t.device} but the current CUDA device is {torch.cuda.current_device()}: every '
                           'tensor of a call must share one device (kernels launch on the current device and its current stream)')
    return C.c_void_p(t.data_ptr())


def texel_ptr(p):
    """Device pointer of a factor tensor [1,C,...] that must be dense in channels-last order."""
    if not p.is_cuda or p.dtype != torch.float32:
        raise RuntimeError('ffb200: factor tensors must be fp32 CUDA tensors')
    fmt = torch.channels_last if p.dim() == 4 else torch.channels_last_3d
    if not (p.is_contiguous(memory_format=fmt) or p.shape[1] == 1 and p.is_contiguous()):
        raise RuntimeError('ffb200: factor tensor is not channels-last contiguous')
    return p.data_ptr()


def i32p(t, allow_none=True):
    return ptr(t, torch.int32, allow_none)


# ---- optional per-section device timing (bench.py's roofline): CUDA events on the launching stream ----------
_sections = None


def profile_begin():
    global _sections
    _sections = {}


def profile_end():
    """-> {name: (total_ms, calls)}; synchronises."""
    global _sections
    torch.cuda.synchronize()
    out = {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in _sections.items()}
    _sections = None
    return out


class section:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _sections is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if _sections is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _sections.setdefault(self.name, []).append((self.e0, e1))
        return False
