"""Autograd-aware wrappers around the libffb200.so entry points.

Each class below is the counterpart of a block of ATen ops in the reference (cited per class).  All
compute is in the CUDA library; torch supplies device memory, the current stream and the autograd
graph so that unmodified caller code (`loss.backward()`, `torch.optim.Adam`) keeps working.
"""
import ctypes as C

import torch

from . import native as nv

_ACT = {'none': 0, 'relu': 1, 'sigmoid': 2}


def _dev_check(t):
    if not t.is_cuda:
        raise RuntimeError('ffb200 ops need CUDA tensors: there is no CPU path (the oracle under oracle/ is test-only)')


_host_cache = {}


def host_list(t):
    """t.tolist() for small, rarely-changing device tensors (aabb, stepSize, freq_bands) without a device->host sync on
    every call: cached by (storage pointer, version counter, shape).  The cache keeps the tensor alive, so its address
    cannot be recycled for another tensor while the entry exists; in-place writes bump the version counter."""
    if not torch.is_tensor(t):
        return t
    if not t.is_cuda:
        return t.tolist()
    key = (t.data_ptr(), t._version, tuple(t.shape))
    v = _host_cache.get(key)
    if v is None:
        if len(_host_cache) > 256:
            _host_cache.clear()
        v = _host_cache[key] = (t, t.tolist())
    return v[1]


# Optional pre-allocated gradient storage (train.TrainStep): {parameter data_ptr: zeroed tensor with the parameter's
# shape and strides}.  While set, the backward passes accumulate into these instead of allocating zeros_like tensors —
# valid only when every parameter is used by exactly one forward call per backward (TrainStep guarantees that).
_grad_arena = None


def set_grad_arena(mapping):
    global _grad_arena
    _grad_arena = mapping


def _grad_like(t):
    if _grad_arena is not None:
        g = _grad_arena.get(t.data_ptr())
        if g is not None:
            return g
    return torch.zeros_like(t)   # preserve_format keeps the channels-last strides


# Two-phase field backward (train.TrainStep under data parallelism): while `_field_bwd_split` is a dict, FieldQuery.backward
# scatters only into the tensors whose data_ptr is NOT in split['late'] and stashes the launch arguments; the caller runs
# `field_bwd_deferred()` later (a second CUDA graph) for the rest.  The early gradients — the fine basis levels, 15 of the
# 21 MB at nerf.yaml — can then be all-reduced while the late scatter (coefficients + coarse levels) is still running.
_field_bwd_split = None


def set_field_bwd_split(split):
    global _field_bwd_split
    _field_bwd_split = split


def field_bwd_deferred():
    """Launch the stashed second phase of the field backward (no-op when nothing was deferred)."""
    sp = _field_bwd_split
    if not sp or not sp.get('stash'):
        return
    for plan, x, n, n_dev, gf, gc, coeff, basis, arr in sp['stash']:
        with nv.section('field_bwd'):
            nv.check(nv.lib().ffb_field_query_bwd_saved(plan.handle, nv.ptr(x), C.c_int64(n), nv.i32p(n_dev), nv.ptr(gf, allow_none=True),
                                                        nv.ptr(gc, allow_none=True), nv.ptr(coeff, allow_none=True),
                                                        nv.ptr(basis, allow_none=True), arr, nv.stream()))


def _empty(shape, like, dtype=torch.float32):
    return torch.empty(shape, device=like.device, dtype=dtype)


# --------------------------------------------------------------------------------------------------
# Field plan: host-side description of get_coeff / get_basis (FactorFields.py:425-516) as gather ops.
# --------------------------------------------------------------------------------------------------
class FieldPlan:
    """Owns an ffb_field_t handle.  `tensors[i]` is the factor tensor op i samples (None for none)."""

    def __init__(self, desc, tensors, perm, width):
        self.desc, self.tensors, self.perm, self.width = desc, tensors, perm, width
        self.handle = C.c_void_p()
        nv.check(nv.lib().ffb_field_create(C.byref(desc), C.byref(self.handle)))
        self.ptrs = [t.data_ptr() if t is not None else 0 for t in tensors]
        self.fast = nv.lib().ffb_field_fast_eligible(self.handle) == 1   # specialised grid x grid kernels apply
        # kernels whose backward scatters from rows saved by the training forward (basis row: blocked for the grid x grid kernels,
        # row-major for the vm kernels)
        # (measured for the vm kernels too — ffb_field_planes_bwd_saved — and left off: writing the permuted basis row costs the
        # forward +0.16 ms at the -vm bench shape and the backward gains nothing over re-gathering L2-resident texels)
        self.saves_rows = self.fast
        # the vm kernels' backward reads the coefficient row the forward returns anyway (a coalesced stream instead of re-gathering
        # the three lines); the plane texels are re-gathered
        self.keeps_coeff = (not self.fast) and nv.lib().ffb_field_planes_eligible(self.handle) == 1

    def stale(self):
        return any((t.data_ptr() if t is not None else 0) != p for t, p in zip(self.tensors, self.ptrs))

    def __del__(self):
        try:
            if self.handle:
                nv.lib().ffb_field_destroy(self.handle)
        except Exception:
            pass


class PlanBuilder:
    def __init__(self, xdim, in_dim, aabb, mapping, freq):
        d = nv.FieldDesc()
        d.xdim, d.in_dim = xdim, in_dim
        lo, hi = host_list(aabb)
        for k in range(len(lo)):
            d.aabb_min[k], d.aabb_max[k] = lo[k], hi[k]
        d.mapping = nv.MAP_IDS[mapping]
        fl = host_list(freq) if freq is not None else []
        if len(fl) > nv.MAX_FREQ:
            raise RuntimeError(f'ffb200: at most {nv.MAX_FREQ} frequency bands are supported')
        d.n_freq = len(fl)
        for i, f in enumerate(fl):
            d.freq[i] = f
        self.d = d
        self.tensors = []
        self.perm = None

    def op(self, tensor, src, space, level=0, align=False, border=False, nearest=False, cst=(0.0, 0.0, 0.0)):
        """tensor [1, C, *spatial] channels-last; src: per grid_sample coordinate (x=W first) the x column or -1."""
        d = self.d
        if d.n_ops >= nv.MAX_OPS:
            raise RuntimeError('ffb200: too many gather ops in one field')
        o = d.ops[d.n_ops]
        o.data = nv.texel_ptr(tensor)
        o.grad = 0
        o.C = tensor.shape[1]
        spatial = list(tensor.shape[2:])[::-1]  # W, H, (D)
        o.nd = len(spatial)
        if len(src) != o.nd:
            raise RuntimeError('ffb200: gather op coordinate count does not match the tensor rank')
        for k in range(o.nd):
            o.size[k], o.src[k], o.cst[k] = spatial[k], src[k], cst[k]
        o.space, o.level, o.align_corners, o.border, o.nearest = space, level, int(align), int(border), int(nearest)
        self.tensors.append(tensor)
        d.n_ops += 1
        return d.n_ops - 1

    def term(self, which, ops, col):
        d = self.d
        n = d.n_cterms if which == 'c' else d.n_bterms
        if n >= nv.MAX_TERMS:
            raise RuntimeError('ffb200: too many terms in one field')
        t = (d.cterms if which == 'c' else d.bterms)[n]
        t.n_ops = len(ops)
        for i, o in enumerate(ops):
            t.op[i] = o
        t.col = col
        if which == 'c':
            d.n_cterms += 1
        else:
            d.n_bterms += 1

    def finish(self, coeff_width, basis_width, basis_is_x=False, perm=None, device=None):
        d = self.d
        d.coeff_width, d.basis_width, d.basis_is_x = coeff_width, basis_width, int(basis_is_x)
        if perm is not None:
            self.perm = torch.as_tensor(perm, dtype=torch.int32, device=device).contiguous()
            d.basis_perm = self.perm.data_ptr()
        else:
            d.basis_perm = 0
        return FieldPlan(d, self.tensors, self.perm, basis_width if basis_width > 0 else coeff_width)


class FieldQuery(torch.autograd.Function):
    """get_coding / get_coeff / get_basis (FactorFields.py:425-533) + the autograd of F.grid_sample w.r.t. the
    factor tensors (the reference never needs d/dx)."""

    @staticmethod
    @nv.on_device_of(2)
    def forward(ctx, plan, x, n_dev, *tensors):
        """n_dev: None, or a 1-element int32 CUDA tensor holding the live row count (x is then a capacity-sized buffer)."""
        _dev_check(x)
        x = x.contiguous().float()
        n = x.shape[0]
        feats = _empty((n, plan.width), x)
        coeff = _empty((n, plan.width), x)
        # training: also keep the basis row, so the backward pass scatters without re-gathering
        train = plan.saves_rows and any(ctx.needs_input_grad[3:])
        basis = _empty(((n + 31) // 32 * 32, plan.width), x) if train else None     # private, blocked by 32 rows (field_fast.cu: blk_idx)
        if n > 0:
            with nv.section('field_fwd'):
                if train:
                    nv.check(nv.lib().ffb_field_query_fwd_train(plan.handle, nv.ptr(x), C.c_int64(n), nv.i32p(n_dev), nv.ptr(feats),
                                                                nv.ptr(coeff), nv.ptr(basis), nv.stream()))
                else:
                    nv.check(nv.lib().ffb_field_query_fwd(plan.handle, nv.ptr(x), C.c_int64(n), nv.i32p(n_dev), nv.ptr(feats), nv.ptr(coeff),
                                                          nv.stream()))
        ctx.plan, ctx.n_dev, ctx.train = plan, n_dev, train
        ctx.keep_coeff = (not train) and plan.keeps_coeff and any(ctx.needs_input_grad[3:])
        ctx.set_materialize_grads(False)     # an unused output (the coefficient row in the render path) arrives as None, not as zeros
        if train:
            ctx.save_for_backward(x, coeff, basis)
        elif ctx.keep_coeff:
            ctx.save_for_backward(x, coeff)
        else:
            ctx.save_for_backward(x)
        return feats, coeff

    @staticmethod
    def backward(ctx, g_feats, g_coeff):
        plan = ctx.plan
        x = ctx.saved_tensors[0]
        coeff, basis = (ctx.saved_tensors[1], ctx.saved_tensors[2]) if ctx.train else (None, None)
        if ctx.keep_coeff:
            coeff = ctx.saved_tensors[1]
        n = x.shape[0]
        if g_feats is None and g_coeff is None:
            return (None, None, None, *[None] * len(plan.tensors))
        # one gradient tensor per distinct factor tensor (an op list may reference a tensor once only)
        grads = []
        arr = (C.c_void_p * nv.MAX_OPS)()
        for i, t in enumerate(plan.tensors):
            if ctx.needs_input_grad[3 + i]:
                g = _grad_like(t)
                grads.append(g)
                arr[i] = g.data_ptr()
            else:
                grads.append(None)
                arr[i] = 0
        if n > 0 and any(g is not None for g in grads):
            gf = g_feats.contiguous() if g_feats is not None else None
            gc = g_coeff.contiguous() if g_coeff is not None else None
            sp = _field_bwd_split
            if sp is not None and plan.fast and ctx.train:
                late = (C.c_void_p * nv.MAX_OPS)()
                n_late = 0
                for i, t in enumerate(plan.tensors):
                    if arr[i] and t.data_ptr() in sp['late']:
                        late[i], arr[i] = arr[i], 0
                        n_late += 1
                if n_late:
                    sp.setdefault('stash', []).append((plan, x, n, ctx.n_dev, gf, gc, coeff, basis, late))
            with nv.section('field_bwd'):
                nv.check(nv.lib().ffb_field_query_bwd_saved(plan.handle, nv.ptr(x), C.c_int64(n), nv.i32p(ctx.n_dev),
                                                            nv.ptr(gf, allow_none=True), nv.ptr(gc, allow_none=True),
                                                            nv.ptr(coeff, allow_none=True), nv.ptr(basis, allow_none=True), arr, nv.stream()))
        return (None, None, None, *grads)


@nv.on_device_of(0)
def grid_mapping(positions, freq_bands, aabb, basis_mapping='sawtooth'):
    """FactorFields.py:11-33.  positions [..., d] -> [..., d, F] ([..., d, 2F] for 'trigonometric')."""
    _dev_check(positions)
    shp = positions.shape
    d = shp[-1]
    x = positions.reshape(-1, d).contiguous().float()
    F = freq_bands.numel()
    Fo = 2 * F if basis_mapping == 'trigonometric' else F
    out = _empty((x.shape[0], d, Fo), x)
    lo = (C.c_float * 3)(*host_list(aabb)[0])
    hi = (C.c_float * 3)(*host_list(aabb)[1])
    fr = (C.c_float * F)(*host_list(freq_bands))
    nv.check(nv.lib().ffb_grid_mapping(nv.ptr(x), C.c_int64(x.shape[0]), d, lo, hi, fr, F, nv.MAP_IDS[basis_mapping], nv.ptr(out),
                                       nv.stream()))
    return out.reshape(*shp, Fo)


# --------------------------------------------------------------------------------------------------
# MLPs
# --------------------------------------------------------------------------------------------------
def _linear_fwd(x, W, b, act, n_dev=None, terms=3):
    """terms: bf16 parts per operand on the tensor-core path (3: fp32-class, 2: ~5e-6 relative)."""
    n, K = x.shape
    M = W.shape[0]
    y = _empty((n, M), x)
    if n > 0:
        nv.check(nv.lib().ffb_linear_fwd_ex(nv.ptr(x), nv.ptr(W), nv.ptr(b, allow_none=True), nv.ptr(y), C.c_int64(n), nv.i32p(n_dev), K, M,
                                            act, terms, nv.stream()))
    return y


def _mlp_backward(acts, layers, acts_kind, g_out, need_input_grad, param_needs, n_dev=None):
    """acts[l] = input of layer l, acts[l+1] = its (activated) output.  -> (g_input | None, [gW, gb, ...])"""
    lib = nv.lib()
    g = g_out.contiguous()
    n = g.shape[0]
    grads = [None] * len(layers)
    for l in range(len(layers) - 1, -1, -1):
        W, b = layers[l]
        M, K = W.shape
        act = acts_kind[l]
        y = acts[l + 1]
        gW = _grad_like(W) if param_needs[l][0] else None
        gb = _grad_like(b) if (b is not None and param_needs[l][1]) else None
        want_gx = l > 0 or need_input_grad
        if M <= 8 and n > 0:
            # colour / density heads: one exact-fp32 pass for gx, gW and gb
            gx = _empty((n, K), g) if want_gx else None
            nv.check(lib.ffb_linear_bwd_skinny(nv.ptr(g), nv.ptr(y, allow_none=(act == 0)), act, nv.ptr(acts[l]), nv.ptr(W),
                                               nv.ptr(gx, allow_none=True), nv.ptr(gW, allow_none=True), nv.ptr(gb, allow_none=True),
                                               C.c_int64(n), nv.i32p(n_dev), K, M, nv.stream()))
            grads[l] = (gW, gb)
            g = gx
            continue
        if n > 0 and gW is not None:
            nv.check(lib.ffb_linear_bwd_weight_act(nv.ptr(g), nv.ptr(y, allow_none=(act == 0)), act, nv.ptr(acts[l]), nv.ptr(gW),
                                                   nv.ptr(gb, allow_none=True), C.c_int64(n), nv.i32p(n_dev), K, M, nv.stream()))
        grads[l] = (gW, gb)
        if want_gx:
            gx = _empty((n, K), g)
            if n > 0:
                nv.check(lib.ffb_linear_bwd_input(nv.ptr(g), nv.ptr(y, allow_none=(act == 0)), nv.ptr(W), nv.ptr(gx), C.c_int64(n),
                                                  nv.i32p(n_dev), K, M, act, nv.stream()))
            g = gx
        else:
            g = None
    return g, grads


def _split_params(params, has_bias):
    layers, i = [], 0
    for hb in has_bias:
        W = params[i]
        b = params[i + 1] if hb else None
        i += 2 if hb else 1
        layers.append((W, b))
    return layers


# Sparse hand-off of the compositor's gradient to linear_mat's backward.  d(loss)/d(linear_mat output) is dense only in column 0
# (density); columns 1.. are non-zero for the shaded samples alone (~15 % at nerf.yaml).  With SPARSE_FEAT_GRAD on — train.TrainStep
# switches it on around the graph it owns, where linear_mat's output feeds RenderComposite and nothing else — RenderComposite.backward
# returns an UNWRITTEN dense-shaped tensor and parks (density gradient [Nv], slot map [Nv], compact rows [Na, ld]) under its
# data_ptr; MLPFunction.backward picks them up and runs ffb_mlp2p_bwd_sparse.  Saves writing and re-reading [Nv, 32] floats
# (2 x 126 MB per step).  `sparse_grads_pending()` lets the owner check that every parked gradient was consumed.
SPARSE_FEAT_GRAD = False
_sparse_grads = {}


def sparse_grads_pending():
    n = len(_sparse_grads)
    _sparse_grads.clear()
    return n


class MLPFunction(torch.autograd.Function):
    """MLPMixer.forward (FactorFields.py:144-159): optional PE concat, Linear+ReLU ..., bias-free last layer."""

    @staticmethod
    @nv.on_device_of(1)
    def forward(ctx, x, pe, has_bias, n_dev, *params):
        _dev_check(x)
        x = x.contiguous().float()
        layers = _split_params(params, has_bias)
        n, D = x.shape
        h = x
        ctx.n_dev = n_dev
        ctx.fused = False
        if (pe == 0 and len(layers) == 2 and layers[0][1] is not None and layers[1][1] is None and (n >= 1024 or n_dev is not None)
                and nv.lib().ffb_mlp2_eligible(D, layers[0][0].shape[0], layers[1][0].shape[0]) == 1):
            # whole MLP in one tcgen05 kernel; the hidden activation is recomputed by the backward kernel
            (W1, b1), (W2, _) = layers
            y = _empty((n, W2.shape[0]), x)
            need_bwd = any(ctx.needs_input_grad)
            relu_bits = torch.empty((n, W1.shape[0] // 16), device=x.device, dtype=torch.int16) if need_bwd else None
            with nv.section('mlp_fwd'):
                nv.check(nv.lib().ffb_mlp2_fwd(nv.ptr(x), nv.ptr(W1), nv.ptr(b1), nv.ptr(W2), nv.ptr(y), nv.ptr(relu_bits, torch.int16, True),
                                               C.c_int64(n), nv.i32p(n_dev), D, W1.shape[0], W2.shape[0], nv.stream()))
            ctx.fused, ctx.has_bias = True, has_bias
            if need_bwd:
                ctx.save_for_backward(x, *params, relu_bits)
            return y
        if pe > 0:
            h = _empty((n, D + 2 * D * pe), x)
            if n > 0:
                nv.check(nv.lib().ffb_pe_concat_fwd(nv.ptr(x), nv.ptr(h), C.c_int64(n), nv.i32p(n_dev), D, pe, nv.stream()))
        acts = [h]
        kinds = []
        with nv.section('mlp_fwd'):
            for l, (W, b) in enumerate(layers):
                act = 1 if l != len(layers) - 1 else 0
                kinds.append(act)
                h = _linear_fwd(h, W, b, act, n_dev)
                acts.append(h)
        ctx.pe, ctx.has_bias, ctx.kinds = pe, has_bias, kinds
        ctx.save_for_backward(x, *acts, *params)
        ctx.n_acts = len(acts)
        return h

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        x = saved[0]
        if ctx.fused:
            W1, b1, W2, relu_bits = saved[1:5]
            needs = ctx.needs_input_grad
            n, D = x.shape
            gx = _empty((n, D), x) if needs[0] else None
            gW1 = _grad_like(W1) if needs[4] else None
            gb1 = _grad_like(b1) if needs[5] else None
            gW2 = _grad_like(W2) if needs[6] else None
            sp = _sparse_grads.pop(g.data_ptr(), None) if _sparse_grads else None
            if sp is not None:
                g0, slot, rows, shape = sp
                if shape != tuple(g.shape) or nv.lib().ffb_mlp2_pipelined_eligible(D, W1.shape[0], W2.shape[0]) != 1:
                    raise RuntimeError('ffb200: a sparse compositor gradient reached an MLP that cannot consume it')
                with nv.section('mlp_bwd'):
                    nv.check(nv.lib().ffb_mlp2p_bwd_sparse(nv.ptr(x), nv.ptr(g0), nv.i32p(slot), nv.ptr(rows), nv.ptr(W1), nv.ptr(b1), nv.ptr(W2),
                                                           nv.ptr(relu_bits, torch.int16), nv.ptr(gx, allow_none=True),
                                                           nv.ptr(gW1, allow_none=True), nv.ptr(gb1, allow_none=True), nv.ptr(gW2, allow_none=True),
                                                           C.c_int64(n), nv.i32p(ctx.n_dev), D, W1.shape[0], W2.shape[0], nv.stream()))
                return (gx, None, None, None, gW1, gb1, gW2)
            with nv.section('mlp_bwd'):
                nv.check(nv.lib().ffb_mlp2_bwd(nv.ptr(x), nv.ptr(g.contiguous()), nv.ptr(W1), nv.ptr(b1), nv.ptr(W2), nv.ptr(relu_bits, torch.int16),
                                               nv.ptr(gx, allow_none=True),
                                               nv.ptr(gW1, allow_none=True), nv.ptr(gb1, allow_none=True), nv.ptr(gW2, allow_none=True),
                                               C.c_int64(n), nv.i32p(ctx.n_dev), D, W1.shape[0], W2.shape[0], nv.stream()))
            return (gx, None, None, None, gW1, gb1, gW2)
        acts = list(saved[1:1 + ctx.n_acts])
        params = saved[1 + ctx.n_acts:]
        layers = _split_params(params, ctx.has_bias)
        needs = ctx.needs_input_grad[4:]
        pn, i = [], 0
        for hb in ctx.has_bias:
            pn.append((needs[i], needs[i + 1] if hb else False))
            i += 2 if hb else 1
        with nv.section('mlp_bwd'):
            gin, grads = _mlp_backward(acts, layers, ctx.kinds, g, ctx.needs_input_grad[0], pn, ctx.n_dev)
        gx = None
        if ctx.needs_input_grad[0]:
            if ctx.pe > 0:
                n, D = x.shape
                gx = _empty((n, D), x)
                if n > 0:
                    nv.check(nv.lib().ffb_pe_concat_bwd(nv.ptr(x), nv.ptr(gin), nv.ptr(gx), C.c_int64(n), nv.i32p(ctx.n_dev), D, ctx.pe,
                                                        nv.stream()))
            else:
                gx = gin
        flat = []
        for (gW, gb), hb in zip(grads, ctx.has_bias):
            flat.append(gW)
            if hb:
                flat.append(gb)
        return (gx, None, None, None, *flat)


# --------------------------------------------------------------------------------------------------
# Sampling + compaction (no gradient)
# --------------------------------------------------------------------------------------------------
SAMPLE_BOUNDED, SAMPLE_NDC, SAMPLE_UNBOUND = 0, 1, 2
# bf16 parts per operand of the appearance MLP forward (3: ~3e-7 relative.  With 2 parts (~5e-6) enough ReLU decisions flip
# against the reference that hidden-layer weight gradients differ by 5e-4 — measured, scratch/diag_relu_flip.py)
APPEARANCE_TERMS = 3


def make_sampler_desc(aabb, step_size, n_samples, alpha=None, alpha_thres=0.5, mode=SAMPLE_BOUNDED, z_table=None, bg_len=0.0,
                      alpha_outside=False):
    """mode / z_table: FFB_SAMPLE_* of include/ffb200.h; z_table is the device [n_samples] interpx row of
    sample_point_ndc / sample_point_unbound (the descriptor keeps a reference so the buffer outlives the launches)."""
    d = nv.SamplerDesc()
    d.mode, d.bg_len, d.alpha_outside = int(mode), float(bg_len), int(bool(alpha_outside))
    if mode != SAMPLE_BOUNDED:
        if z_table is None or not z_table.is_cuda or z_table.dtype != torch.float32 or z_table.numel() != int(n_samples):
            raise RuntimeError('NDC / unbounded sampling needs a CUDA float32 z_table with n_samples entries')
        d._keep = z_table.contiguous()
        d.z_table = d._keep.data_ptr()
    else:
        d.z_table = 0
    lo, hi = host_list(aabb)
    for k in range(3):
        d.aabb_min[k], d.aabb_max[k] = lo[k], hi[k]
    d.step_size = float(host_list(step_size))
    d.n_samples = int(n_samples)
    if alpha is not None:
        d.alpha_volume = alpha.volume_u8.data_ptr()
        D, H, W = alpha.volume_u8.shape
        d.alpha_size[0], d.alpha_size[1], d.alpha_size[2] = W, H, D
        alo, ainv = host_list(alpha.aabb)[0], host_list(alpha.invgridSize)
        for k in range(3):
            d.alpha_aabb_min[k], d.alpha_inv_size[k] = alo[k], ainv[k]
        d.alpha_thres = alpha_thres
    else:
        d.alpha_volume = 0
    return d


@nv.on_device_of(0)
def exclusive_scan(counts):
    R = counts.shape[0]
    out = _empty((R + 1,), counts, torch.int32)
    nv.check(nv.lib().ffb_exclusive_scan_i32(nv.i32p(counts), nv.i32p(out), C.c_int64(R), nv.stream()))
    return out


@torch.no_grad()
@nv.on_device_of(1)
def sample_compact(desc, rays, jitter, lazy=False):
    """-> dict(xyz [Nv,3], ray_id, sample_id, z, dist, offsets [R+1], n_valid, n_dev).
    lazy=False: one host read of Nv, exact-size buffers (n_dev None).  lazy=True: no host round trip; buffers have the
    capacity R*S and every consumer reads the live count from n_dev (= offsets[R], a device int32)."""
    _dev_check(rays)
    lib = nv.lib()
    rays = rays.contiguous().float()
    R = rays.shape[0]
    counts = _empty((R,), rays, torch.int32)
    tmin = _empty((R,), rays)
    jp = nv.ptr(jitter, allow_none=True)
    sec = nv.section('sample')
    sec.__enter__()
    nv.check(lib.ffb_sample_count(C.byref(desc), nv.ptr(rays), jp, C.c_int64(R), nv.i32p(counts), nv.ptr(tmin), nv.stream()))
    offsets = exclusive_scan(counts)
    n_dev = None
    if lazy:
        nvld, n_dev = R * desc.n_samples, offsets[R:R + 1]
    else:
        nvld = int(offsets[R].item())
    xyz = _empty((nvld, 3), rays)
    ray_id = _empty((nvld,), rays, torch.int32)
    sample_id = _empty((nvld,), rays, torch.int32)
    z = _empty((nvld,), rays)
    dist = _empty((nvld,), rays)
    if nvld > 0:
        nv.check(lib.ffb_sample_fill(C.byref(desc), nv.ptr(rays), jp, nv.ptr(tmin), nv.i32p(offsets), C.c_int64(R), C.c_int64(nvld),
                                     nv.ptr(xyz), nv.i32p(ray_id), nv.i32p(sample_id), nv.ptr(z), nv.ptr(dist), nv.stream()))
    sec.__exit__()
    return dict(xyz=xyz, ray_id=ray_id, sample_id=sample_id, z=z, dist=dist, offsets=offsets, counts=counts,
                n_valid=(offsets[R] if lazy else nvld), n_dev=n_dev, rays=rays)


@torch.no_grad()
@nv.on_device_of(1)
def sample_dense(desc, rays, jitter, want_z=True, want_pts=False):
    _dev_check(rays)
    rays = rays.contiguous().float()
    R, S = rays.shape[0], desc.n_samples
    mask = _empty((R, S), rays, torch.uint8)
    z = _empty((R, S), rays) if want_z else None
    pts = _empty((R, S, 3), rays) if want_pts else None
    nv.check(nv.lib().ffb_sample_dense(C.byref(desc), nv.ptr(rays), nv.ptr(jitter, allow_none=True), C.c_int64(R),
                                       nv.ptr(mask, torch.uint8), nv.ptr(z, allow_none=True), nv.ptr(pts, allow_none=True), nv.stream()))
    if want_pts:
        return mask.bool(), z, pts
    return mask.bool(), z


# --------------------------------------------------------------------------------------------------
# Composite + appearance MLP
# --------------------------------------------------------------------------------------------------
def make_composite_desc(density_shift, fea2denseAct, distance_scale, weight_thres, white_bg, white_bg_dev=None):
    """white_bg_dev: optional 1-element int32 CUDA tensor that overrides `white_bg` on the device (the descriptor keeps a
    reference so the buffer outlives the launches)."""
    d = nv.CompositeDesc()
    d._keep = white_bg_dev
    d.white_bg_dev = nv.i32p(white_bg_dev).value if white_bg_dev is not None else None
    d.density_shift = float(density_shift)
    d.softplus = 1 if fea2denseAct == 'softplus' else 0
    d.distance_scale = float(distance_scale)
    d.weight_thres = float(weight_thres)
    d.white_bg = int(bool(white_bg))
    return d


class RenderComposite(torch.autograd.Function):
    """FactorFields.forward from `sigma[ray_valid] = ...` to `rgb_map.clamp` (FactorFields.py:876-896): density
    activation, raw2alpha, weight-threshold compaction, MLPRender_Fea on the shaded samples, accumulation."""

    @staticmethod
    @nv.on_device_of(1)
    def forward(ctx, feat, samp, cdesc, view_pe, fea_pe, has_bias, *params):
        lib = nv.lib()
        feat = feat.contiguous()
        Nv, ld = feat.shape
        rays, offsets = samp['rays'], samp['offsets']
        R = rays.shape[0]
        if Nv == 0:   # every ray misses the box / the alpha mask: background only (FactorFields.py:887-893 with weight == 0)
            if cdesc.white_bg_dev:
                rgb0 = (cdesc._keep != 0).to(torch.float32).expand(R, 3).contiguous()
            else:
                rgb0 = torch.full((R, 3), 1.0 if cdesc.white_bg else 0.0, device=feat.device)
            ctx.Na, ctx.empty, ctx.has_bias = 0, True, has_bias
            ctx.save_for_backward(feat, *params)
            outs = (rgb0, torch.zeros(R, device=feat.device), torch.zeros(R, device=feat.device),
                    _empty((0,), feat), _empty((0,), feat, torch.int32), torch.tensor(0))
            ctx.mark_non_differentiable(*outs[1:])
            return outs
        ctx.empty = False
        sigma, trans, weight = _empty((Nv,), feat), _empty((Nv,), feat), _empty((Nv,), feat)
        app_counts = _empty((R,), feat, torch.int32)
        sec = nv.section('composite_fwd')
        sec.__enter__()
        nv.check(lib.ffb_composite_weights(C.byref(cdesc), nv.ptr(feat), ld, nv.ptr(samp['dist']), nv.i32p(offsets), C.c_int64(R),
                                           nv.ptr(sigma), nv.ptr(trans), nv.ptr(weight), nv.i32p(app_counts), nv.stream()))
        app_offsets = exclusive_scan(app_counts)
        lazy = samp.get('n_dev') is not None
        a_dev = app_offsets[R:R + 1] if lazy else None
        Na = Nv if lazy else int(app_offsets[R].item())     # lazy: capacity; the live count stays on the device
        app_idx = _empty((Na,), feat, torch.int32)
        sec.__exit__()
        layers = _split_params(params, has_bias)
        Cf = ld - 1
        Win = 3 + Cf + 6 * view_pe + 2 * fea_pe * Cf
        acts, kinds = [], []
        app_slot = None
        if Na > 0:
            if SPARSE_FEAT_GRAD and ctx.needs_input_grad[0] and ld % 4 == 0:
                app_slot = _empty((Nv,), feat, torch.int32)      # inverse of app_idx, for the sparse gradient hand-off
            nv.check(lib.ffb_composite_app_fill_ex(nv.ptr(weight), C.c_float(cdesc.weight_thres), nv.i32p(offsets), nv.i32p(app_offsets),
                                                   C.c_int64(R), nv.i32p(app_idx), nv.i32p(app_slot), nv.stream()))
            ws_bytes = 0
            if len(layers) == 3 and tuple(has_bias) == (True, True, False) and layers[2][0].shape[0] == 3 \
                    and layers[1][0].shape[0] == layers[1][0].shape[1] == layers[0][0].shape[0] and layers[0][0].shape[1] == Win \
                    and (lazy or Na >= 1024):
                ws_bytes = int(lib.ffb_rgbmlp_workspace_bytes(Cf, layers[0][0].shape[0], view_pe, fea_pe))
            if ws_bytes > 0:
                # whole appearance MLP in one tcgen05 kernel (mlp_rgb.cu).  For the backward kernel it leaves the ReLU
                # decision bits and bf16 operand streams of x / h1 / h2 (already in the layout the TMA engine feeds to the MMAs)
                (W1, b1), (W2, b2), (W3, _) = layers
                need_bwd = any(ctx.needs_input_grad)
                ws = torch.empty(ws_bytes, device=feat.device, dtype=torch.uint8)
                h = _empty((Na, 3), feat)
                bits = sx = sh1 = sh2 = None
                if need_bwd:
                    bits = torch.empty((Na, 16), device=feat.device, dtype=torch.int16)
                    nbx, nbh = (int(lib.ffb_rgbmlp_stream_bytes(Cf, view_pe, fea_pe, C.c_int64(Na), w)) for w in (0, 1))
                    sx = torch.empty(nbx, device=feat.device, dtype=torch.uint8)
                    sh1 = torch.empty(nbh, device=feat.device, dtype=torch.uint8)
                    sh2 = torch.empty(nbh, device=feat.device, dtype=torch.uint8)
                vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
                with nv.section('rgbmlp_fwd'):
                    nv.check(lib.ffb_rgbmlp_pack(nv.ptr(W1), nv.ptr(b1), nv.ptr(W2), nv.ptr(W3), vp(ws), Cf, view_pe, fea_pe, nv.stream()))
                    nv.check(lib.ffb_rgbmlp_fwd(nv.ptr(feat), ld, nv.ptr(rays), nv.i32p(samp['ray_id']), nv.i32p(app_idx), vp(ws), nv.ptr(b2),
                                                nv.ptr(h), vp(bits), None, None, None, vp(sx), vp(sh1), vp(sh2), C.c_int64(Na),
                                                nv.i32p(a_dev), Cf, view_pe, fea_pe, nv.stream()))
                acts, kinds = [ws, bits, sx, sh1, sh2], 'fused'
            else:
                inp = _empty((Na, Win), feat)
                nv.check(lib.ffb_render_input_fwd(nv.ptr(feat), ld, nv.ptr(rays), nv.i32p(samp['ray_id']), nv.i32p(app_idx), nv.ptr(inp),
                                                  C.c_int64(Na), nv.i32p(a_dev), Cf, view_pe, fea_pe, nv.stream()))
                h = inp
                acts.append(h)
                with nv.section('rgbmlp_fwd'):
                    for l, (W, b) in enumerate(layers):
                        act = 1 if l != len(layers) - 1 else 2   # ReLU ... sigmoid (FactorFields.py:197-202)
                        kinds.append(act)
                        h = _linear_fwd(h, W, b, act, a_dev, terms=APPEARANCE_TERMS)
                        acts.append(h)
            rgb = h
        else:
            rgb = _empty((0, 3), feat)
        rgb_map, pre_clamp = _empty((R, 3), feat), _empty((R, 3), feat)
        acc, depth = _empty((R,), feat), _empty((R,), feat)
        nv.check(lib.ffb_composite_accum(C.byref(cdesc), nv.ptr(weight), nv.ptr(samp['z']), nv.ptr(rgb, allow_none=True) if Na > 0 else None, nv.i32p(offsets),
                                         nv.i32p(app_offsets), C.c_int64(R), nv.ptr(rgb_map), nv.ptr(pre_clamp), nv.ptr(acc),
                                         nv.ptr(depth), nv.stream()))
        ctx.cdesc, ctx.samp, ctx.has_bias, ctx.kinds = cdesc, samp, has_bias, kinds
        ctx.app_slot = app_slot if kinds == 'fused' else None
        ctx.view_pe, ctx.fea_pe, ctx.n_acts, ctx.Na, ctx.a_dev = view_pe, fea_pe, len(acts), Na, a_dev
        ctx.save_for_backward(feat, sigma, trans, weight, rgb, app_offsets, app_idx, pre_clamp, *acts, *params)
        n_app = app_offsets[R] if lazy else torch.tensor(Na)
        ctx.mark_non_differentiable(depth, acc, weight, app_idx, n_app)
        return rgb_map, depth, acc, weight, app_idx, n_app

    @staticmethod
    def backward(ctx, g_rgb_map, *_unused):
        lib = nv.lib()
        saved = ctx.saved_tensors
        if ctx.empty:
            return (torch.zeros_like(saved[0]), None, None, None, None, None, *[torch.zeros_like(p) for p in saved[1:]])
        feat, sigma, trans, weight, rgb, app_offsets, app_idx, pre_clamp = saved[:8]
        acts = list(saved[8:8 + ctx.n_acts])
        params = saved[8 + ctx.n_acts:]
        samp, cdesc = ctx.samp, ctx.cdesc
        Nv, ld = feat.shape
        R = samp['rays'].shape[0]
        Na = ctx.Na
        sparse = ctx.app_slot is not None and Na > 0 and Nv > 0
        zero_rest = 1 if (ld % 4 == 0 and Nv > 0 and not sparse) else 0     # the composite kernel writes whole gradient rows: no memset of [Nv, ld]
        # sparse hand-off (SPARSE_FEAT_GRAD): g_feat stays unwritten, the density gradient goes to a compact [Nv] vector
        g_feat = torch.empty_like(feat) if (zero_rest or sparse) else torch.zeros_like(feat)
        g0 = _empty((Nv,), feat) if sparse else None
        g_rgb = _empty((Na, 3), feat)
        g_rgb_map = g_rgb_map.contiguous()
        if Nv > 0:
            with nv.section('composite_bwd'):
                nv.check(lib.ffb_composite_bwd(C.byref(cdesc), nv.ptr(g_rgb_map), nv.ptr(pre_clamp), nv.ptr(feat), ld,
                                               nv.ptr(samp['dist']), nv.ptr(sigma), nv.ptr(trans), nv.ptr(weight),
                                               nv.ptr(rgb) if Na > 0 else None,
                                               nv.i32p(samp['offsets']), nv.i32p(app_offsets), C.c_int64(R), nv.ptr(g_rgb) if Na > 0 else None,
                                               nv.ptr(g0 if sparse else g_feat), 1 if sparse else ld, zero_rest, nv.stream()))
        layers = _split_params(params, ctx.has_bias)
        needs = ctx.needs_input_grad[6:]
        pn, i = [], 0
        for hb in ctx.has_bias:
            pn.append((needs[i], needs[i + 1] if hb else False))
            i += 2 if hb else 1
        flat = []
        if Na > 0:
            Cf = ld - 1
            if ctx.kinds == 'fused':
                ws, bits, sx, sh1, sh2 = acts
                (W1, b1), (W2, b2), (W3, _) = layers
                grads = [(_grad_like(W) if nw else None, _grad_like(b) if (b is not None and nb) else None)
                         for (W, b), (nw, nb) in zip(layers, pn)]
                ld_gin = (W1.shape[1] + 3) // 4 * 4          # 16-byte aligned rows: the kernel writes them with vector stores
                g_in = _empty((Na, ld_gin), feat)
                vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
                with nv.section('rgbmlp_bwd'):
                    nv.check(lib.ffb_rgbmlp_bwd(nv.ptr(g_rgb), nv.ptr(rgb), vp(bits), vp(sx), vp(sh1), vp(sh2), vp(ws), nv.ptr(W3), nv.ptr(g_in),
                                                ld_gin, vp(grads[0][0]), vp(grads[0][1]), vp(grads[1][0]), vp(grads[1][1]), vp(grads[2][0]),
                                                C.c_int64(Na), nv.i32p(ctx.a_dev), Cf, ctx.view_pe, ctx.fea_pe, nv.stream()))
            else:
                with nv.section('rgbmlp_bwd'):
                    g_in, grads = _mlp_backward(acts, layers, ctx.kinds, g_rgb, True, pn, ctx.a_dev)
            if sparse:
                g_app = _empty((Na, ld), feat)
                nv.check(lib.ffb_render_input_bwd_compact(nv.ptr(feat), ld, nv.i32p(app_idx), nv.ptr(g_in), g_in.shape[1], nv.ptr(g_app), ld,
                                                          C.c_int64(Na), nv.i32p(ctx.a_dev), Cf, ctx.view_pe, ctx.fea_pe, nv.stream()))
                _sparse_grads[g_feat.data_ptr()] = (g0, ctx.app_slot, g_app, tuple(g_feat.shape))
            else:
                nv.check(lib.ffb_render_input_bwd(nv.ptr(feat), ld, nv.i32p(app_idx), nv.ptr(g_in), g_in.shape[1], nv.ptr(g_feat), C.c_int64(Na),
                                                  nv.i32p(ctx.a_dev), Cf, ctx.view_pe, ctx.fea_pe, nv.stream()))
            for (gW, gb), hb in zip(grads, ctx.has_bias):
                flat.append(gW)
                if hb:
                    flat.append(gb)
        else:
            for (W, b), hb, (nw, nb) in zip(layers, ctx.has_bias, pn):
                flat.append(_grad_like(W) if nw else None)
                if hb:
                    flat.append(_grad_like(b) if nb else None)
        return (g_feat, None, None, None, None, None, *flat)


class RenderMLP(torch.autograd.Function):
    """MLPRender_Fea.forward (FactorFields.py:188-203) for callers that use the module directly."""

    @staticmethod
    @nv.on_device_of(1)
    def forward(ctx, viewdirs, features, view_pe, fea_pe, has_bias, *params):
        _dev_check(features)
        lib = nv.lib()
        n, Cf = features.shape
        # pack [unused, features] so the shared input-assembly kernel (which skips the density column) can be used
        feat = torch.cat([torch.zeros_like(features[:, :1]), features], -1).contiguous()
        rays = torch.cat([torch.zeros_like(viewdirs), viewdirs], -1).contiguous()
        Win = 3 + Cf + 6 * view_pe + 2 * fea_pe * Cf
        inp = _empty((n, Win), features)
        if n > 0:
            nv.check(lib.ffb_render_input_fwd(nv.ptr(feat), Cf + 1, nv.ptr(rays), None, None, nv.ptr(inp), C.c_int64(n), None, Cf, view_pe,
                                              fea_pe, nv.stream()))
        layers = _split_params(params, has_bias)
        acts, kinds, h = [inp], [], inp
        for l, (W, b) in enumerate(layers):
            act = 1 if l != len(layers) - 1 else 2
            kinds.append(act)
            h = _linear_fwd(h, W, b, act, terms=2)
            acts.append(h)
        ctx.has_bias, ctx.kinds, ctx.n_acts, ctx.view_pe, ctx.fea_pe = has_bias, kinds, len(acts), view_pe, fea_pe
        ctx.save_for_backward(feat, *acts, *params)
        return h

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        feat = saved[0]
        acts = list(saved[1:1 + ctx.n_acts])
        params = saved[1 + ctx.n_acts:]
        layers = _split_params(params, ctx.has_bias)
        needs = ctx.needs_input_grad[5:]
        pn, i = [], 0
        for hb in ctx.has_bias:
            pn.append((needs[i], needs[i + 1] if hb else False))
            i += 2 if hb else 1
        g_in, grads = _mlp_backward(acts, layers, ctx.kinds, g, True, pn)
        n, ld = feat.shape
        g_feat = torch.zeros_like(feat)
        if n > 0:
            nv.check(nv.lib().ffb_render_input_bwd(nv.ptr(feat), ld, None, nv.ptr(g_in), 0, nv.ptr(g_feat), C.c_int64(n), None, ld - 1,
                                                   ctx.view_pe, ctx.fea_pe, nv.stream()))
        flat = []
        for (gW, gb), hb in zip(grads, ctx.has_bias):
            flat.append(gW)
            if hb:
                flat.append(gb)
        return (None, g_feat[:, 1:], None, None, None, *flat)


@torch.no_grad()
@nv.on_device_of(1)
def density_alpha(cdesc, feat, length):
    n, ld = feat.shape
    out = _empty((n,), feat)
    if n > 0:
        nv.check(nv.lib().ffb_density_alpha(C.byref(cdesc), nv.ptr(feat.contiguous()), ld, C.c_float(float(length)), C.c_int64(n), None,
                                            nv.ptr(out), nv.stream()))
    return out


@torch.no_grad()
@nv.on_device_of(0)
def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    """In-place fused Adam on the raw storage (any memory format, as long as p/g/m/v share it)."""
    n = p.numel()
    nv.check(nv.lib().ffb_adam_step(C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(m.data_ptr()), C.c_void_p(v.data_ptr()),
                                    C.c_int64(n), C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps), int(step),
                                    C.c_float(grad_scale), nv.stream()))


@torch.no_grad()
@nv.on_device_of(0)
def adam_hyper_advance(lr_d, step_d, hyper_d, beta1, beta2, lr_decay):
    """lr_d [G] float64, step_d [1] int64, hyper_d [G,2] float32 — all on the device."""
    nv.check(nv.lib().ffb_adam_hyper_advance(nv.ptr(lr_d, torch.float64), nv.ptr(step_d, torch.int64), nv.ptr(hyper_d), lr_d.numel(),
                                             C.c_double(beta1), C.c_double(beta2), C.c_double(lr_decay), nv.stream()))


@torch.no_grad()
@nv.on_device_of(0)
def adam_multi(table, chunk_tensor, chunk_start, chunk, hyper_d, beta1, beta2, eps, grad_scale=1.0):
    nv.check(nv.lib().ffb_adam_multi(nv.ptr(table, torch.int64), nv.ptr(chunk_tensor, torch.int32), nv.ptr(chunk_start, torch.int64),
                                     chunk_tensor.numel(), int(chunk), nv.ptr(hyper_d), C.c_float(beta1), C.c_float(beta2), C.c_float(eps),
                                     C.c_float(grad_scale), nv.stream()))


@torch.no_grad()
@nv.on_device_of(0)
def scalar_decay(value_d, factor, out_f32=None):
    """value_d [1] float64 (device) *= factor; out_f32 [1] float32 receives the rounded copy."""
    nv.check(nv.lib().ffb_scalar_decay(nv.ptr(value_d, torch.float64), C.c_double(factor), nv.ptr(out_f32, allow_none=True), nv.stream()))


@torch.no_grad()
@nv.on_device_of(0)
def mse_fwd_bwd(pred, target, g_scale=1.0, loss=None, g_scale_dev=None):
    """-> (loss [1] device tensor, g_pred); g_scale_dev: optional device float32 scalar multiplied into g_pred."""
    pred, target = pred.contiguous(), target.contiguous()
    if loss is None:
        loss = torch.zeros(1, device=pred.device)
    else:
        loss.zero_()
    g = torch.empty_like(pred)
    nv.check(nv.lib().ffb_mse_fwd_bwd(nv.ptr(pred), nv.ptr(target), C.c_int64(pred.numel()), C.c_float(g_scale), nv.ptr(g_scale_dev, allow_none=True),
                                      nv.ptr(loss), nv.ptr(g), nv.stream()))
    return loss, g
