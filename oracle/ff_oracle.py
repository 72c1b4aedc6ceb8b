"""ORACLE — test infrastructure, NOT product code.

CPU (numpy, fp32) restatement of the Factor Fields query-and-render path, i.e. of
`/root/reference/models/FactorFields.py` plus the PyTorch ATen operators that the
reference calls (`F.grid_sample`, `torch.remainder`, `torch.cumprod`, `nn.Linear`,
`F.softplus`).  ATen is a third-party dependency that is not vendored in the
reference (the reference pins torch 1.13.0, README.md:9,16; this container has
torch 2.11.0); its published grid-sampler algorithm is restated in
`grid_sample`/`grid_sample_bwd` below and anchored on the reference's call sites.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` leg may import this module.  The product (`factor-fields_b200/`) must
never import it: the product has no CPU path at all.

PARITY PIN: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the *unmodified reference itself*, imported in
the dev container by `tests/golden/make_golden.py`; the vectors are committed under
`tests/golden/*.npz` and checked by `tests/test_oracle_golden.py`.

Every function cites the reference lines it follows (paths relative to
/root/reference).
"""
import math
import numpy as np

f32 = np.float32


def _f(x):
    return np.asarray(x, dtype=np.float32)


# --------------------------------------------------------------------------------------
# ATen grid sampler (third-party; semantics used at FactorFields.py:105,433,439,446-450,
# 458,489,495,502-508).  Coordinates: pts[..., 0] <-> W (last axis), [...,1] <-> H, [...,2] <-> D.
# --------------------------------------------------------------------------------------
def _unnormalize(coord, size, align_corners):
    if align_corners:
        return ((coord + f32(1)) / f32(2)) * f32(size - 1)
    return ((coord + f32(1)) * f32(size) - f32(1)) / f32(2)


def _source_index(coord, size, align_corners, padding):
    c = _unnormalize(coord, size, align_corners)
    if padding == 'border':
        c = np.minimum(f32(size - 1), np.maximum(c, f32(0)))
    return c.astype(np.float32)


def _corners(inp_spatial, pts, mode, align_corners, padding):
    """Yield (index tuple (slowest..fastest), weight[N], inbounds[N]) in ATen's accumulation order."""
    nd = len(inp_spatial)
    sizes = inp_spatial[::-1]  # W, H, (D)
    src = [_source_index(pts[:, k], sizes[k], align_corners, padding) for k in range(nd)]
    if mode == 'nearest':
        idx = [np.rint(s).astype(np.int64) for s in src]
        inb = np.ones(pts.shape[0], bool)
        for k in range(nd):
            inb &= (idx[k] >= 0) & (idx[k] < sizes[k])
        return [(tuple(idx[::-1]), np.ones(pts.shape[0], np.float32), inb)]
    lo = [np.floor(s) for s in src]
    w0 = [(l + f32(1)) - s for l, s in zip(lo, src)]  # weight of the low corner: (ix_hi - ix)
    w1 = [s - l for l, s in zip(lo, src)]             # weight of the high corner: (ix - ix_lo)
    lo_i = [l.astype(np.int64) for l in lo]
    out = []
    # ATen order: x fastest (nw, ne, sw, se), then z (top four, bottom four)
    for corner in range(1 << nd):
        bits = [(corner >> k) & 1 for k in range(nd)]  # bit0 -> x, bit1 -> y, bit2 -> z
        idx = [lo_i[k] + bits[k] for k in range(nd)]
        w = (w1[0] if bits[0] else w0[0]).astype(np.float32)
        for k in range(1, nd):
            w = (w * (w1[k] if bits[k] else w0[k])).astype(np.float32)
        inb = np.ones(pts.shape[0], bool)
        for k in range(nd):
            inb &= (idx[k] >= 0) & (idx[k] < sizes[k])
        out.append((tuple(idx[::-1]), w, inb))
    return out


def grid_sample(inp, pts, mode='bilinear', align_corners=False, padding='zeros'):
    """inp [C, (D,) H, W] fp32, pts [N, nd] in [-1,1] -> [N, C]."""
    inp = _f(inp)
    pts = _f(pts)
    C = inp.shape[0]
    spatial = inp.shape[1:]
    out = np.zeros((pts.shape[0], C), np.float32)
    for idx, w, inb in _corners(spatial, pts, mode, align_corners, padding):
        safe = tuple(np.where(inb, i, 0) for i in idx)
        vals = inp[(slice(None),) + safe].T  # [N, C]
        contrib = (vals * w[:, None]).astype(np.float32)
        out = (out + np.where(inb[:, None], contrib, f32(0))).astype(np.float32)
    return out


def grid_sample_bwd(inp_shape, pts, gout, mode='bilinear', align_corners=False, padding='zeros'):
    """Gradient of grid_sample w.r.t. `inp` (the only gradient the reference ever needs:
    sample positions never require grad).  gout [N, C] -> [C, (D,) H, W]."""
    pts = _f(pts)
    gout = _f(gout)
    C = inp_shape[0]
    spatial = tuple(inp_shape[1:])
    ginp = np.zeros((int(np.prod(spatial)), C), np.float64)
    strides = np.cumprod((1,) + spatial[::-1][:-1])[::-1]
    for idx, w, inb in _corners(spatial, pts, mode, align_corners, padding):
        if not inb.any():
            continue
        flat = sum(i[inb] * int(s) for i, s in zip(idx, strides))
        np.add.at(ginp, flat, (gout[inb] * w[inb, None]).astype(np.float32))
    return ginp.T.reshape((C,) + spatial).astype(np.float32)


# --------------------------------------------------------------------------------------
# FactorFields.py:11-33
# --------------------------------------------------------------------------------------
def grid_mapping(positions, freq_bands, aabb, basis_mapping='sawtooth'):
    positions = _f(positions)
    freq_bands = _f(freq_bands)
    aabb = _f(aabb)
    aabbSize = np.max(aabb[1] - aabb[0]).astype(np.float32)
    scale = (aabbSize / freq_bands).astype(np.float32)  # [F]
    p = (positions - aabb[0])[..., None].astype(np.float32)
    if basis_mapping == 'triangle':
        local = np.remainder(p, scale).astype(np.float32)
        local_int = np.remainder(np.floor_divide(p, scale), f32(2))
        local = (local / (scale / f32(2)) - f32(1)).astype(np.float32)
        local = np.where(local_int == 1, -local, local)
    elif basis_mapping == 'sawtooth':
        local = np.remainder(p, scale).astype(np.float32)
        local = (local / (scale / f32(2)) - f32(1)).astype(np.float32)
        local = np.clip(local, f32(-1), f32(1))
    elif basis_mapping == 'sinc':
        local = np.sin(p / (scale / f32(np.pi)) - f32(np.pi / 2))
    elif basis_mapping == 'trigonometric':
        local = p / scale * f32(2) * f32(np.pi)
        local = np.concatenate((np.sin(local), np.cos(local)), axis=-1)
    elif basis_mapping == 'x':
        local = p / scale
    else:
        raise ValueError(basis_mapping)
    return local.astype(np.float32)


# FactorFields.py:74-79
def positional_encoding(positions, freqs):
    positions = _f(positions)
    fb = (f32(2) ** np.arange(freqs, dtype=np.float32)).astype(np.float32)
    pts = (positions[..., None] * fb).reshape(positions.shape[:-1] + (freqs * positions.shape[-1],))
    return np.concatenate([np.sin(pts), np.cos(pts)], axis=-1).astype(np.float32)


def positional_encoding_bwd(positions, freqs, g):
    positions = _f(positions)
    fb = (f32(2) ** np.arange(freqs, dtype=np.float32)).astype(np.float32)
    pts = (positions[..., None] * fb)
    D = positions.shape[-1]
    gs = g[..., :D * freqs].reshape(pts.shape)
    gc = g[..., D * freqs:].reshape(pts.shape)
    return ((gs * np.cos(pts) - gc * np.sin(pts)) * fb).sum(-1).astype(np.float32)


# FactorFields.py:82-88
def raw2alpha(sigma, dist):
    sigma = _f(sigma)
    dist = _f(dist)
    alpha = (f32(1) - np.exp(-sigma * dist)).astype(np.float32)
    ones = np.ones_like(alpha[..., :1])
    fac = (f32(1) - alpha + f32(1e-10)).astype(np.float32)
    T = np.cumprod(np.concatenate([ones, fac], -1), -1, dtype=np.float32)
    weights = (alpha * T[..., :-1]).astype(np.float32)
    return alpha, weights, T[..., -1:]


def softplus(x):
    x = _f(x)
    return np.where(x > f32(20), x, np.log1p(np.exp(np.minimum(x, f32(20))))).astype(np.float32)


def sigmoid(x):
    x = _f(x)
    return (f32(1) / (f32(1) + np.exp(-x))).astype(np.float32)


# --------------------------------------------------------------------------------------
# MLPMixer (FactorFields.py:113-159) and MLPRender_Fea (:162-203)
# layers = [(W [out,in], b [out] or None), ...]; ReLU between, last layer bias-free.
# --------------------------------------------------------------------------------------
def mlp_forward(layers, x, pe=0, dropout_mask=None, want_cache=False):
    h = _f(x)
    x0 = h
    if pe > 0:
        h = np.concatenate([h, positional_encoding(h, pe)], axis=-1)
    if dropout_mask is not None:  # F.dropout(p=0.1): mask in {0,1}, scale 1/0.9  (:150-151)
        h = (h * dropout_mask * f32(1.0 / 0.9)).astype(np.float32)
    acts = [h]
    for l, (W, b) in enumerate(layers):
        h = h @ _f(W).T
        if b is not None:
            h = h + _f(b)
        if l != len(layers) - 1:
            h = np.maximum(h, f32(0))
        h = h.astype(np.float32)
        acts.append(h)
    if want_cache:
        return h, (x0, acts, pe, dropout_mask)
    return h


def mlp_backward(layers, cache, gout):
    """-> (g_input (w.r.t. the un-encoded input), [(gW, gb), ...])"""
    x0, acts, pe, dropout_mask = cache
    g = _f(gout)
    grads = [None] * len(layers)
    for l in range(len(layers) - 1, -1, -1):
        W, b = layers[l]
        if l != len(layers) - 1:
            g = np.where(acts[l + 1] > 0, g, f32(0)).astype(np.float32)
        gW = (g.T @ acts[l]).astype(np.float32)
        gb = g.sum(0).astype(np.float32) if b is not None else None
        grads[l] = (gW, gb)
        g = (g @ _f(W)).astype(np.float32)
    if dropout_mask is not None:
        g = (g * dropout_mask * f32(1.0 / 0.9)).astype(np.float32)
    if pe > 0:
        D = x0.shape[-1]
        g = (g[..., :D] + positional_encoding_bwd(x0, pe, g[..., D:])).astype(np.float32)
    return g, grads


def render_mlp_input(viewdirs, features, viewpe, feape):
    indata = [_f(features), _f(viewdirs)]
    if feape > 0:
        indata.append(positional_encoding(features, feape))
    if viewpe > 0:
        indata.append(positional_encoding(viewdirs, viewpe))
    return np.concatenate(indata, axis=-1).astype(np.float32)


def render_mlp_forward(layers, viewdirs, features, viewpe=6, feape=2, want_cache=False):
    h0 = render_mlp_input(viewdirs, features, viewpe, feape)
    h, cache = mlp_forward(layers, h0, want_cache=True)
    rgb = sigmoid(h)
    if want_cache:
        return rgb, (cache, rgb, _f(features), feape)
    return rgb


def render_mlp_backward(layers, cache, g_rgb):
    """-> (g_features, layer grads); view directions never need a gradient."""
    mcache, rgb, features, feape = cache
    g = (_f(g_rgb) * rgb * (f32(1) - rgb)).astype(np.float32)
    g_in, grads = mlp_backward(layers, mcache, g)
    C = features.shape[-1]
    g_feat = g_in[:, :C].copy()
    if feape > 0:
        off = C + 3
        g_feat = g_feat + positional_encoding_bwd(features, feape, g_in[:, off:off + 2 * feape * C])
    return g_feat.astype(np.float32), grads


# --------------------------------------------------------------------------------------
# The field: get_coeff (:425-465), get_basis (:467-516), get_coding (:523-533)
# --------------------------------------------------------------------------------------
class FieldOracle:
    """spec: dict(mode, in_dim, aabb [2,d], coeff_type, basis_type, basis_mapping, coef_mode,
    basis_mode, basis_dims, freq_bands, n_scene, scene_idx).  params: dict(coeffs=[...],
    basises=[...]) in the reference layout [1, C, (D,) H, W]; for 'mlp' factor types the list
    entries are layer lists [(W, b), ...]."""

    matMode = [[0, 1], [0, 2], [1, 2]]
    vecMode = [2, 1, 0]

    def __init__(self, spec, params):
        self.s = dict(spec)
        self.s.setdefault('n_scene', 1)
        self.s.setdefault('scene_idx', 0)
        self.p = params
        self.aabb = _f(self.s['aabb'])
        self.in_dim = int(self.s['in_dim'])
        self.freq_bands = _f(self.s['freq_bands'])

    # :635-637
    def normalize_coord(self, x):
        inv = (f32(2.0) / (self.aabb[1] - self.aabb[0])).astype(np.float32)
        return ((_f(x) - self.aabb[0]) * inv - f32(1)).astype(np.float32)

    def _mode(self, m):
        return 'bilinear' if m in ('bilinear', 'linear') else m

    def _scene_col(self):
        return f32((self.s['scene_idx'] + 0.5) / self.s['n_scene'] * 2 - 1)

    # ---- coefficient -----------------------------------------------------------------
    def _coeff_terms(self, x):
        """List of terms; a term = (out column offset, [ (param_key, pts, mode, align, pad) ... ]) whose
        samples are multiplied channel-wise."""
        ct = self.s['coeff_type']
        pts = self.normalize_coord(x)
        N, dim = pts.shape
        mode = self._mode(self.s['coef_mode'])
        terms = []
        if 'grid' in ct:
            terms.append((0, [(('coeffs', self.s['scene_idx']), pts, mode, False, 'border')]))
        elif 'vec' in ct:
            col = np.full(N, self._scene_col(), np.float32)
            terms.append((0, [(('coeffs', 0), np.stack([col, pts[:, 0]], -1), mode, False, 'border')]))
        elif 'cp' in ct:
            col = np.full(N, self._scene_col(), np.float32)
            terms.append((0, [(('coeffs', i), np.stack([col, pts[:, i]], -1), mode, False, 'border')
                              for i in range(self.in_dim)]))
        elif 'vm' in ct:
            col = np.full(N, self._scene_col(), np.float32)
            off = 0
            for i in range(self.in_dim):
                C = self.p['coeffs'][i].shape[1]
                terms.append((off, [(('coeffs', i), np.stack([col, pts[:, self.vecMode[i]]], -1),
                                     mode, False, 'border')]))
                off += C
        else:
            raise ValueError(ct)
        return terms

    def _eval_terms(self, terms, width):
        N = terms[0][1][0][1].shape[0]
        out = np.zeros((N, width), np.float32)
        for off, ops in terms:
            val = None
            for (key, idx), pts, mode, align, pad in ops:
                v = grid_sample(self.p[key][idx][0], pts, mode, align, pad)
                val = v if val is None else (val * v).astype(np.float32)
            out[:, off:off + val.shape[1]] = val
        return out

    def _terms_width(self, terms):
        w = 0
        for off, ops in terms:
            (key, idx) = ops[0][0]
            w = max(w, off + self.p[key][idx].shape[1])
        return w

    def _bwd_terms(self, terms, g, grads):
        for off, ops in terms:
            vals = [grid_sample(self.p[k][i][0], pts, mode, align, pad) for (k, i), pts, mode, align, pad in ops]
            C = vals[0].shape[1]
            gt = g[:, off:off + C]
            for j, ((k, i), pts, mode, align, pad) in enumerate(ops):
                gj = gt.copy()
                for jj, v in enumerate(vals):
                    if jj != j:
                        gj = (gj * v).astype(np.float32)
                gi = grid_sample_bwd(self.p[k][i][0].shape, pts, gj, mode, align, pad)[None]
                grads[k][i] = gi if grads[k][i] is None else grads[k][i] + gi

    def get_coeff(self, x):
        ct = self.s['coeff_type']
        if 'mlp' in ct:
            return mlp_forward(self.p['coeffs'][self.s['scene_idx']], self.normalize_coord(x), pe=4)
        terms = self._coeff_terms(x)
        return self._eval_terms(terms, self._terms_width(terms))

    # ---- basis -----------------------------------------------------------------------
    def _basis_x(self, x):
        x = _f(x)
        if self.s['mode'] == 'images':
            x = x[..., :-1]
        return x

    def _mapped(self, x):
        """-> xyz [N, in_dim, F]: the reference's `.view(..., -1, in_dim, freq_len)` (:481-482).
        For 'trigonometric' this re-interprets [N, d, 2F] memory as [2N, d, F]."""
        F = len(self.freq_bands)
        m = grid_mapping(x, self.freq_bands, self.aabb[:, :self.in_dim], self.s['basis_mapping'])
        return np.ascontiguousarray(m).reshape(-1, self.in_dim, F)

    def _basis_terms(self, x):
        bt = self.s['basis_type']
        xyz = self._mapped(x)
        F = len(self.freq_bands)
        mode = self._mode(self.s['basis_mode'])
        terms, off = [], 0
        for i in range(F):
            if 'grid' in bt:
                C = self.p['basises'][i].shape[1]
                terms.append((off, [(('basises', i), xyz[..., i], mode, True, 'zeros')]))
                off += C
            elif 'vm' in bt:
                for m in range(self.in_dim):
                    k = i * self.in_dim + m
                    C = self.p['basises'][k].shape[1]
                    pts = np.stack([xyz[:, self.matMode[m][0], i], xyz[:, self.matMode[m][1], i]], -1)
                    terms.append((off, [(('basises', k), pts, 'bilinear', True, 'zeros')]))
                    off += C
            elif 'cp' in bt:
                ops = []
                for a in range(self.in_dim - 1):
                    k = i * (self.in_dim - 1) + a
                    v = xyz[:, a + 1, i]
                    ops.append((('basises', k), np.stack([np.zeros_like(v), v], -1), 'bilinear', True, 'zeros'))
                C = self.p['basises'][i * (self.in_dim - 1)].shape[1]
                terms.append((off, ops))
                off += C
            else:
                raise ValueError(bt)
        return terms, off

    def _vm_perm(self, total):
        """:514-515  basises.view(N, F, -1).permute(0, 2, 1).reshape(N, -1): out[:, perm[q]] = cat[:, q]."""
        F = len(self.freq_bands)
        per = total // F
        q = np.arange(total)
        return (q % per) * F + q // per

    def get_basis(self, x):
        bt = self.s['basis_type']
        x = self._basis_x(x)
        N = x.shape[0]
        F = len(self.freq_bands)
        if 'mlp' in bt:
            xyz = self._mapped(x)
            outs = [mlp_forward(self.p['basises'][i], xyz[..., i].reshape(-1, self.in_dim), pe=4) for i in range(F)]
            return np.concatenate(outs, -1)
        if 'grid' in bt or 'vm' in bt or 'cp' in bt:
            terms, width = self._basis_terms(x)
            cat = self._eval_terms(terms, width)
            if 'vm' in bt:
                out = np.empty_like(cat)
                out[:, self._vm_perm(width)] = cat
                return out
            return cat
        if 'x' in bt:
            xyz = self._mapped(x)
            return np.concatenate([xyz[..., i].reshape(N, -1) for i in range(F)], -1).astype(np.float32)
        raise ValueError(bt)

    def get_coding(self, x):
        ct, bt = self.s['coeff_type'], self.s['basis_type']
        if ct != 'none' and bt != 'none':
            c = self.get_coeff(x)
            b = self.get_basis(x)
            return (b * c).astype(np.float32), c
        if ct != 'none':
            c = self.get_coeff(x)
            return c, c
        b = self.get_basis(x)
        return b, b

    def get_coding_bwd(self, x, g_feats, g_coeff=None):
        """Gradient of sum(feats * g_feats) (+ sum(coeff * g_coeff): get_coding's second output, FactorFields.py:527) w.r.t.
        the factor tensors (grid / vec / cp / vm types).  -> dict(coeffs=[...], basises=[...]) in the reference layout."""
        ct, bt = self.s['coeff_type'], self.s['basis_type']
        g_feats = _f(g_feats)
        grads = {'coeffs': [None] * len(self.p.get('coeffs', [])),
                 'basises': [None] * len(self.p.get('basises', []))}
        have_c = ct != 'none'
        have_b = bt != 'none'
        c = self.get_coeff(x) if have_c else None
        b = self.get_basis(x) if have_b else None
        gc = (g_feats * b).astype(np.float32) if (have_c and have_b) else g_feats
        gb = (g_feats * c).astype(np.float32) if (have_c and have_b) else g_feats
        if g_coeff is not None and have_c and have_b:
            gc = (gc + _f(g_coeff)).astype(np.float32)
        if have_c and 'mlp' in ct:
            i = self.s['scene_idx']
            _, cache = mlp_forward(self.p['coeffs'][i], self.normalize_coord(x), pe=4, want_cache=True)
            grads['coeffs'][i] = mlp_backward(self.p['coeffs'][i], cache, gc)[1]
        elif have_c:
            self._bwd_terms(self._coeff_terms(x), gc, grads)
        if have_b and 'mlp' in bt:
            xyz = self._mapped(self._basis_x(x))
            off = 0
            for i in range(len(self.freq_bands)):
                xi = xyz[..., i].reshape(-1, self.in_dim)
                y, cache = mlp_forward(self.p['basises'][i], xi, pe=4, want_cache=True)
                grads['basises'][i] = mlp_backward(self.p['basises'][i], cache, gb[:, off:off + y.shape[1]])[1]
                off += y.shape[1]
        if have_b and ('grid' in bt or 'vm' in bt or 'cp' in bt):
            terms, width = self._basis_terms(self._basis_x(x))
            if 'vm' in bt:
                gb = gb[:, self._vm_perm(width)]
            self._bwd_terms(terms, gb, grads)
        for k in grads:
            for i, g in enumerate(grads[k]):
                if g is None and not isinstance(self.p[k][i], list):
                    grads[k][i] = np.zeros_like(self.p[k][i])
        return grads


# --------------------------------------------------------------------------------------
# Ray sampling (:586-602), alpha mask (:91-110), composite / forward (:843-898)
# --------------------------------------------------------------------------------------
def sample_point(aabb, stepSize, rays_o, rays_d, N_samples, jitter=None):
    """jitter: [R] uniform numbers (the reference draws them with torch.rand_like on the CPU
    generator, :595) or None for is_train=False.  -> pts [R,S,3], z [R,S], in-box mask [R,S]."""
    aabb = _f(aabb)
    o = _f(rays_o)
    d = _f(rays_d)
    vec = np.where(d == 0, f32(1e-6), d).astype(np.float32)
    rate_a = ((aabb[1] - o) / vec).astype(np.float32)
    rate_b = ((aabb[0] - o) / vec).astype(np.float32)
    t_min = np.clip(np.minimum(rate_a, rate_b).max(-1), f32(0.05), f32(1e3)).astype(np.float32)
    rng = np.arange(N_samples, dtype=np.float32)[None]
    if jitter is not None:
        rng = (np.repeat(rng, o.shape[0], 0) + _f(jitter)[:, None]).astype(np.float32)
    step = (f32(stepSize) * rng).astype(np.float32)
    interpx = (t_min[:, None] + step).astype(np.float32)
    pts = (o[:, None, :] + (d[:, None, :] * interpx[..., None]).astype(np.float32)).astype(np.float32)
    out = ((aabb[0] > pts) | (pts > aabb[1])).any(-1)
    return pts, interpx, ~out


def ndc_interpx(near, far, N_samples, uniform=None):
    """:577-580 — linspace(near, far, N) (+ uniform [N] * (far-near)/N when training), fp32."""
    interpx = _linspace(near, far, N_samples)
    if uniform is not None:
        interpx = (interpx + (_f(uniform) * f32((far - near) / N_samples)).astype(np.float32)).astype(np.float32)
    return interpx


def _linspace(start, end, steps):
    """torch.linspace for float32 on the CPU (ATen cpu/RangeFactoriesKernel.cpp, third-party): step = (end-start)/
    (steps-1) in fp32; element i is fma(step, i, start) for i < steps//2 and fma(-step, steps-1-i, end) otherwise
    (the compiled kernel contracts the multiply-add; step*i is exact in fp64, so fp64 evaluation + one rounding
    reproduces it — checked here against torch.linspace on 300 random (start, end, steps))."""
    if steps == 1:
        return np.array([start], np.float32)
    start, end = f32(start), f32(end)
    step = f32((end - start) / f32(steps - 1))
    i = np.arange(steps)
    lo = (np.float64(start) + np.float64(step) * i).astype(np.float32)
    hi = (np.float64(end) - np.float64(step) * (steps - 1 - i)).astype(np.float32)
    return np.where(i < steps // 2, lo, hi).astype(np.float32)


def sample_point_ndc(aabb, interpx, rays_o, rays_d):
    """:575-584 with interpx [S] given (ndc_interpx) -> pts [R,S,3], z [1,S], in-box mask [R,S]."""
    aabb, o, d = _f(aabb), _f(rays_o), _f(rays_d)
    z = _f(interpx)[None]
    pts = (o[:, None, :] + (d[:, None, :] * z[..., None]).astype(np.float32)).astype(np.float32)
    out = ((aabb[0] > pts) | (pts > aabb[1])).any(-1)
    return pts, z, ~out


def unbound_interpx(N_samples, uniform=None):
    """:607-623 — bin edges linspace(0,2,Ni+1) and 2/linspace(1,1/16,No+1); a uniform point (training) or the
    midpoint (evaluation) of every bin."""
    Ni, No = 3 * N_samples // 4, N_samples // 4
    bi = _linspace(0, 2, Ni + 1)
    bo = (f32(2) / _linspace(1, 1 / 16, No + 1)).astype(np.float32)
    if uniform is not None:
        u = _f(uniform)
        one = f32(1)
        a = ((bi[1:] * u[:Ni]).astype(np.float32) + (bi[:-1] * (one - u[:Ni]).astype(np.float32)).astype(np.float32)).astype(np.float32)
        b = ((bo[1:] * u[Ni:]).astype(np.float32) + (bo[:-1] * (one - u[Ni:]).astype(np.float32)).astype(np.float32)).astype(np.float32)
    else:
        a = ((bi[1:] + bi[:-1]).astype(np.float32) * f32(0.5)).astype(np.float32)
        b = ((bo[1:] + bo[:-1]).astype(np.float32) * f32(0.5)).astype(np.float32)
    return np.concatenate([a, b]).astype(np.float32)


def sample_point_unbound(bg_len, interpx, rays_o, rays_d):
    """:625-633 -> contracted pts [R,S,3], z [1,S], inner mask [R,S]."""
    o, d = _f(rays_o), _f(rays_d)
    z = _f(interpx)[None]
    pts = (o[:, None, :] + (d[:, None, :] * z[..., None]).astype(np.float32)).astype(np.float32)
    norm = np.abs(pts).max(-1, keepdims=True)
    inner = norm <= 1
    with np.errstate(divide='ignore', invalid='ignore'):
        # `self.bg_len / norm` is Tensor.__rtruediv__ = norm.reciprocal() * bg_len
        s = (f32(1 + bg_len) - ((f32(1) / norm).astype(np.float32) * f32(bg_len)).astype(np.float32)).astype(np.float32)
        con = ((pts / norm).astype(np.float32) * s).astype(np.float32)
    pts = np.where(inner, pts, con).astype(np.float32)
    return pts, z, inner[..., 0]


def sample_alpha(alpha_volume, aabb, xyz):
    """AlphaGridMask.sample_alpha (:103-110); alpha_volume [D,H,W] float 0/1."""
    aabb = _f(aabb)
    inv = (f32(1.0) / (aabb[1] - aabb[0]) * f32(2)).astype(np.float32)
    u = ((_f(xyz) - aabb[0]) * inv - f32(1)).astype(np.float32)
    return grid_sample(_f(alpha_volume)[None], u, 'bilinear', True, 'zeros')[:, 0]


def update_render_params(aabb, gridSize, step_ratio):
    """:693-699 -> (stepSize fp32, nSamples)"""
    aabb = _f(aabb)
    aabbSize = aabb[1] - aabb[0]
    units = (aabbSize / (np.asarray(gridSize, np.int64) - 1).astype(np.float32)).astype(np.float32)
    stepSize = f32(np.mean(units, dtype=np.float32) * f32(step_ratio))
    diag = np.sqrt(np.sum(np.square(aabbSize), dtype=np.float32))
    return stepSize, int(f32(diag) / stepSize) + 1


class RenderOracle:
    """forward() of the reference (:843-898) for bounded, non-NDC scenes, dense [R,S] layout like
    the reference, plus its backward.  rspec: dict(aabb, stepSize, distance_scale, density_shift,
    fea2denseAct, rayMarch_weight_thres, view_pe, fea_pe); mlps: dict(linear_mat=[(W,b)..],
    renderModule=[(W,b)..]); alpha = None | dict(volume [D,H,W], aabb)."""

    def __init__(self, field, rspec, mlps, alpha=None):
        self.field, self.r, self.mlps, self.alpha = field, dict(rspec), mlps, alpha

    def basis2density(self, f):
        x = (_f(f) + f32(self.r['density_shift'])).astype(np.float32)
        return softplus(x) if self.r.get('fea2denseAct', 'softplus') == 'softplus' else np.maximum(x, f32(0))

    def basis2density_grad(self, f):
        x = (_f(f) + f32(self.r['density_shift'])).astype(np.float32)
        if self.r.get('fea2denseAct', 'softplus') == 'softplus':
            return np.where(x > f32(20), f32(1), sigmoid(x)).astype(np.float32)
        return (x > 0).astype(np.float32)

    def forward(self, rays, N_samples, jitter=None, white_bg=True, want_cache=False, mode='bounded'):
        """mode 'bounded' (jitter [R], :858-861), 'ndc' (jitter = uniform [S], rspec near_far; :851-857) or
        'unbound' (jitter = uniform [3S//4 + S//4], rspec bg_len; :847-850)."""
        rays = _f(rays)
        o, viewdirs = rays[:, :3], rays[:, 3:6]
        if mode == 'unbound':
            pts, z, inner = sample_point_unbound(self.r['bg_len'], unbound_interpx(N_samples, jitter), o, viewdirs)
            dists = np.concatenate([z[:, 1:] - z[:, :-1], z[:, -1:] - z[:, -2:-1]], -1).astype(np.float32)
            R, S = o.shape[0], z.shape[1]
            valid = np.ones((R, S), bool)
            dists = np.broadcast_to(dists, (R, S))
            z = np.broadcast_to(z, (R, S))
        elif mode == 'ndc':
            near, far = self.r['near_far']
            pts, z, inner = sample_point_ndc(self.r['aabb'], ndc_interpx(near, far, N_samples, jitter), o, viewdirs)
            dists = np.concatenate([z[:, 1:] - z[:, :-1], np.zeros_like(z[:, :1])], -1).astype(np.float32)
            norm = np.sqrt((viewdirs * viewdirs).sum(-1, keepdims=True, dtype=np.float32)).astype(np.float32)
            dists = (dists * norm).astype(np.float32)
            viewdirs = (viewdirs / norm).astype(np.float32)
            R, S = dists.shape
            z = np.broadcast_to(z, (R, S))
            valid = inner.copy()
        else:
            pts, z, inner = sample_point(self.r['aabb'], self.r['stepSize'], o, viewdirs, N_samples, jitter)
            R, S = z.shape
            dists = np.concatenate([z[:, 1:] - z[:, :-1], np.zeros_like(z[:, :1])], -1).astype(np.float32)
            valid = inner.copy()
        if self.alpha is not None:
            a = sample_alpha(self.alpha['volume'], self.alpha['aabb'], pts[inner]) > f32(0.5)
            valid[inner] = a
        sigma = np.zeros((R, S), np.float32)
        rgb = np.zeros((R, S, 3), np.float32)
        feat = None
        lm_cache = None
        if valid.any():
            feats, coeffs = self.field.get_coding(pts[valid])
            feat, lm_cache = mlp_forward(self.mlps['linear_mat'], feats, want_cache=True)
            sigma[valid] = self.basis2density(feat[:, 0])
        else:
            feats = np.zeros((0, 1), np.float32)
            coeffs = np.zeros((1, 1), np.float32)
        delta = (dists * f32(self.r['distance_scale'])).astype(np.float32)
        alpha, weight, _ = raw2alpha(sigma, delta)
        app = weight > f32(self.r['rayMarch_weight_thres'])
        valid_new = valid & app
        app_c = valid_new[valid]
        rm_cache = None
        if app_c.any():
            vd = np.broadcast_to(viewdirs[:, None, :], pts.shape)[valid_new]
            c, rm_cache = render_mlp_forward(self.mlps['renderModule'], vd, feat[app_c, 1:],
                                             self.r['view_pe'], self.r['fea_pe'], want_cache=True)
            rgb[valid_new] = c
        acc = weight.sum(-1, dtype=np.float32)
        rgb_map = (weight[..., None] * rgb).sum(-2, dtype=np.float32)
        if white_bg:
            rgb_map = (rgb_map + (f32(1) - acc[:, None])).astype(np.float32)
        pre_clamp = rgb_map
        rgb_map = np.clip(rgb_map, f32(0), f32(1))
        depth = (weight * z).sum(-1, dtype=np.float32)
        out = dict(rgb_map=rgb_map, depth_map=depth, coeffs=coeffs, ray_valid=valid, weight=weight,
                   sigma=sigma, app_mask=valid_new, feats=feats, z=z, pts=pts)
        if want_cache:
            out['cache'] = dict(pts=pts, valid=valid, valid_new=valid_new, app_c=app_c, feat=feat, lm=lm_cache,
                                rm=rm_cache, alpha=alpha, weight=weight, rgb=rgb, delta=delta, sigma=sigma,
                                pre_clamp=pre_clamp, white_bg=white_bg)
        return out

    def backward(self, cache, g_rgb_map):
        """-> dict(coeffs, basises, linear_mat=[(gW,gb)], renderModule=[(gW,gb)])"""
        c = cache
        g = _f(g_rgb_map).copy()
        g = np.where((c['pre_clamp'] >= 0) & (c['pre_clamp'] <= 1), g, f32(0)).astype(np.float32)
        w, rgb = c['weight'], c['rgb']
        g_rgb = (w[..., None] * g[:, None, :]).astype(np.float32)           # [R,S,3]
        g_w = (rgb * g[:, None, :]).sum(-1)
        if c['white_bg']:
            g_w = g_w - g.sum(-1)[:, None]
        g_w = g_w.astype(np.float64)
        # w_i = a_i T_i ; T_i = prod_{j<i} (1 - a_j + 1e-10)
        a = c['alpha'].astype(np.float64)
        fac = (1.0 - a + 1e-10)
        T = np.cumprod(np.concatenate([np.ones_like(a[:, :1]), fac], -1), -1)[:, :-1]
        gw_w = g_w * (a * T)
        suffix = np.cumsum(gw_w[:, ::-1], -1)[:, ::-1] - gw_w               # sum_{k>i}
        g_alpha = g_w * T - suffix / fac
        g_sigma = (g_alpha * np.exp(-c['sigma'].astype(np.float64) * c['delta']) * c['delta']).astype(np.float32)
        grads = dict(linear_mat=None, renderModule=None)
        valid = c['valid']
        feat = c['feat']
        g_feat = np.zeros_like(feat)
        g_feat[:, 0] = g_sigma[valid] * self.basis2density_grad(feat[:, 0])
        if c['rm'] is not None:
            g_app, grads['renderModule'] = render_mlp_backward(self.mlps['renderModule'], c['rm'],
                                                               g_rgb[c['valid_new']])
            g_feat[c['app_c'], 1:] += g_app
        g_feats, grads['linear_mat'] = mlp_backward(self.mlps['linear_mat'], c['lm'], g_feat)
        grads.update(self.field.get_coding_bwd(c['pts'][valid], g_feats))
        return grads


# --------------------------------------------------------------------------------------
# Alpha-mask maintenance (:710-841)
# --------------------------------------------------------------------------------------
def compute_alpha(render, xyz, length=1.0):
    """FactorFields.compute_alpha (:710-727) for a RenderOracle `render`: alpha = 1 - exp(-basis2density(linear_mat(feats)[..., 0])
    * length) at the points that pass the alpha mask (`> 0`, :713), 0 elsewhere."""
    xyz = _f(xyz)
    keep = np.ones(xyz.shape[0], bool)
    if render.alpha is not None:
        keep = sample_alpha(render.alpha['volume'], render.alpha['aabb'], xyz) > f32(0)
    sigma = np.zeros(xyz.shape[0], np.float32)
    if keep.any():
        feats, _ = render.field.get_coding(xyz[keep])
        feat = mlp_forward(render.mlps['linear_mat'], feats)
        sigma[keep] = render.basis2density(feat[:, 0])
    return (f32(1) - np.exp(-(sigma * f32(length)).astype(np.float32))).astype(np.float32)


def dense_lattice(aabb, gridSize):
    """Voxel-centre lattice of getDenseAlpha (:733-745) -> dense_xyz [D,H,W,3] (already transposed like :746) and the step size."""
    aabb = _f(aabb)
    gs = np.asarray(gridSize, np.int64)
    units = ((aabb[1] - aabb[0]) / (gs - 1).astype(np.float32)).astype(np.float32)
    half = (f32(1.0) / (gs - 1).astype(np.float32) * f32(0.5)).astype(np.float32)
    axes = [_linspace(half[k], f32(1) - half[k], int(gs[k])) for k in range(3)]
    sx, sy, sz = np.meshgrid(*axes, indexing='ij')
    samples = np.stack([sx, sy, sz], -1).astype(np.float32)
    dense = (aabb[0] * (f32(1) - samples) + aabb[1] * samples).astype(np.float32)
    return np.ascontiguousarray(dense.transpose(2, 1, 0, 3)), f32(np.mean(units, dtype=np.float32))


def dense_alpha(render, aabb, gridSize):
    """getDenseAlpha (:730-755) with times = 1 (no jitter): alpha [D,H,W] on the lattice, length = stepSize * distance_scale."""
    dense, step = dense_lattice(aabb, gridSize)
    length = f32(step * f32(render.r['distance_scale']))
    out = np.stack([compute_alpha(render, dense[i].reshape(-1, 3), length).reshape(dense.shape[1], dense.shape[2]) for i in range(dense.shape[0])])
    return out.astype(np.float32), dense


def filter_rays_mask(render, rays, N_samples, bbox_only=False):
    """filtering_rays (:811-841): which rays are kept.  bbox_only: slab test t_max > t_min; else: any of the N_samples
    points (eval sampling, in-box or not) has alpha-mask value > 0 (:832-833)."""
    rays = _f(rays)
    o, d = rays[:, :3], rays[:, 3:6]
    aabb = _f(render.r['aabb'])
    if bbox_only:
        vec = np.where(d == 0, f32(1e-6), d).astype(np.float32)
        ra, rb = ((aabb[1] - o) / vec).astype(np.float32), ((aabb[0] - o) / vec).astype(np.float32)
        return np.maximum(ra, rb).min(-1) > np.minimum(ra, rb).max(-1)
    pts, _, _ = sample_point(aabb, render.r['stepSize'], o, d, N_samples, None)
    a = sample_alpha(render.alpha['volume'], render.alpha['aabb'], pts.reshape(-1, 3)).reshape(pts.shape[:2])
    return (a > f32(0)).any(-1)


def mse_loss_and_grad(rgb_map, target):
    """train_per_scene.py:158: loss = mean((rgb_map - rgb_train)**2)"""
    d = (_f(rgb_map) - _f(target)).astype(np.float32)
    return f32(np.mean(d * d, dtype=np.float32)), (d * f32(2.0 / d.size)).astype(np.float32)


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.99, eps=1e-8):
    """torch.optim.Adam (train_per_scene.py:132, betas=(0.9,0.99)), single-tensor form, fp32."""
    m = (m * f32(beta1) + g * f32(1 - beta1)).astype(np.float32)
    v = (v * f32(beta2) + (g * g) * f32(1 - beta2)).astype(np.float32)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (np.sqrt(v) / f32(math.sqrt(bc2)) + f32(eps)).astype(np.float32)
    p = (p - f32(lr / bc1) * (m / denom)).astype(np.float32)
    return p, m, v
