"""Train-step glue (train_per_scene.py:124-171): fused Adam over the model's parameter groups with the reference's
per-step multiplicative lr decay, and the flat gradient bucket used for the data-parallel all-reduce."""
import os
import sys

import torch

from . import ops


def shard_slice(n, rank, world):
    """Contiguous slice of a global batch of n rays (or sample points) owned by `rank`: disjoint, ordered, covering
    [0, n); the first n % world ranks get one extra element (every rank draws the same global permutation
    and takes its own slice)."""
    base, rem = divmod(int(n), int(world))
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def _dist_world(group=None):
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(group), torch.distributed.get_world_size(group)
    return 0, 1


def gather_rows(local, n_total, group=None):
    """All ranks hold consecutive row shards (shard_slice order) of a [n_total, ...] tensor; returns the whole tensor on
    every rank.  Shards may differ by one row, so they are padded to the largest for the fixed-size all-gather."""
    rank, world = _dist_world(group)
    if world == 1:
        return local
    per = (int(n_total) + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    torch.distributed.all_gather(out, pad, group=group)
    return torch.cat([o[:shard_slice(n_total, r, world).stop - shard_slice(n_total, r, world).start] for r, o in enumerate(out)])


@torch.no_grad()
def render_sharded(rays, model, chunk=65536, N_samples=-1, white_bg=True, ndc_ray=False, group=None, render_fn=None):
    """Evaluation render sharded by ray: rank r renders rows shard_slice(N, r, world) of `rays` [N, 6] (the
    same tensor on every rank) and the rgb / depth maps are all-gathered, so every rank returns the full (rgb_map [N,3],
    depth_map [N]) — what renderer.py:29-98 computes per test image on one device."""
    from .renderer import render_ray
    rank, world = _dist_world(group)
    N = rays.shape[0]
    sl = shard_slice(N, rank, world)
    fn = render_fn or (lambda r: render_ray(r, model, chunk=chunk, N_samples=N_samples, ndc_ray=ndc_ray, white_bg=white_bg,
                                            is_train=False, device=model.device))
    rgb, depth = fn(rays[sl])
    return gather_rows(rgb, N, group), gather_rows(depth, N, group)


def arena_ranges(bucket, keep):
    """Maximal contiguous [start, stop) ranges of the flat gradient arena covering the parameters with keep[i] True."""
    out = []
    for i, k in enumerate(keep):
        if not k:
            continue
        a, b = bucket.offsets[i], bucket.offsets[i + 1]
        if out and out[-1][1] == a:
            out[-1][1] = b
        else:
            out.append([a, b])
    return [tuple(r) for r in out]


def image_set_shard(n_img, rank, world):
    """Image-set mode sharded by image: rank r owns the coefficient slabs of images shard_slice(n_img, r, world).
    Returns (first image, number of local images).  A rank builds its model with aabb[1][-1] = local count, samples pixels
    of its own images only and feeds z - first_image; no coefficient ever crosses ranks (the reference's 3-D coefficient
    tensor blends neighbouring slabs with weight ~1e-6 at z = image + 0.5, FactorFields.py:287,433 — inside the 1e-4 bar)."""
    sl = shard_slice(n_img, rank, world)
    return sl.start, sl.stop - sl.start


class SymmArena:
    """The gradient arena in symmetric memory (torch.distributed._symmetric_memory: every rank maps every peer's buffer and,
    where the NVSwitch fabric offers it, one multicast address), reduced by OUR kernel (csrc/allreduce.cu: multimem.ld_reduce /
    multimem.st through the switch, or peer loads / stores over NVLink) instead of an NCCL call — a plain stream-ordered
    launch, so the data-parallel step is again ONE CUDA graph.  Raises if symmetric memory cannot be set up on this system
    (TrainStep then falls back to NCCL and says so)."""

    BLOCKS = int(os.environ.get('FFB_ALLREDUCE_BLOCKS', '96'))

    def __init__(self, n_floats, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        dist = torch.distributed
        group = group if group is not None else dist.group.WORLD
        n_floats = (int(n_floats) + 3) // 4 * 4
        self.flat = symm_mem.empty(n_floats, dtype=torch.float32, device=device)
        self.flat.zero_()
        self.hdl = symm_mem.rendezvous(self.flat, group)
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        if self.world * 4 > int(self.hdl.signal_pad_size):
            raise RuntimeError('signal pad too small for the all-reduce barrier slots')
        off = int(getattr(self.hdl, 'offset', 0))
        ptrs = [int(p) + off for p in self.hdl.buffer_ptrs]
        if ptrs[self.rank] != self.flat.data_ptr():
            raise RuntimeError('symmetric buffer address does not match the arena tensor')
        self.peers = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.pads = torch.tensor([int(p) for p in self.hdl.signal_pad_ptrs], dtype=torch.int64, device=device)
        mc = int(self.hdl.multicast_ptr or 0)         # 0 when the fabric / driver offers no multicast object for this buffer
        # measured on B200 (tests/dist/dist_check_allreduce.py, 21.4 MB): 2 GPUs — peer loads / stores 52 us, multimem 71 us, NCCL 61 us;
        # 8 GPUs — multimem 72 us (32 CTAs), peer path 91 us, NCCL 115 us.  So: the switch reduction from 3 ranks up.
        nvls = os.environ.get('FFB_ALLREDUCE_NVLS')
        use_nvls = (self.world > 2) if nvls is None else (nvls != '0')
        self.multicast = (mc + off) if (mc and use_nvls) else 0
        if 'FFB_ALLREDUCE_BLOCKS' not in os.environ:
            self.BLOCKS = 32 if self.multicast else 148
        self.epoch = torch.zeros(4, dtype=torch.int32, device=device)
        dist.barrier(group)

    def all_reduce(self):
        from . import native as nv
        import ctypes as C
        nv.check(nv.lib().ffb_allreduce_symm(C.c_void_p(self.flat.data_ptr()), C.c_void_p(self.peers.data_ptr()), C.c_uint64(self.multicast),
                                             C.c_void_p(self.pads.data_ptr()), C.c_void_p(self.epoch.data_ptr()), self.rank, self.world,
                                             C.c_int64(self.flat.numel()), self.BLOCKS, nv.stream()))


class GradBucket:
    """Flat fp32 buffer holding every gradient back to back (parameter storage order), so that one all-reduce per
    step covers grids + MLPs.  Device-agnostic torch code (tested with gloo on CPU)."""

    def __init__(self, params, flat_alloc=None):
        self.params = [p for p in params]
        self.sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for n in self.sizes:      # every slice starts on a 256-byte boundary (the scatter kernels use 16-byte vector reductions)
            self.offsets.append((self.offsets[-1] + n + 63) // 64 * 64)
        dev = self.params[0].device
        self.flat = flat_alloc(self.offsets[-1]) if flat_alloc is not None else torch.zeros(self.offsets[-1], device=dev, dtype=torch.float32)

    def view(self, i):
        """Gradient slice of parameter i in the parameter's own memory order (as_strided over the flat storage)."""
        p = self.params[i]
        return torch.as_strided(self.flat, p.shape, p.stride(), self.offsets[i])

    def pack(self, grads):
        for i, g in enumerate(grads):
            v = self.view(i)
            if g is None:
                v.zero_()
            else:
                v.copy_(g)

    def all_reduce(self, group=None):
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
            torch.distributed.all_reduce(self.flat, op=torch.distributed.ReduceOp.SUM, group=group)
            return torch.distributed.get_world_size(group)
        return 1


class FusedAdam:
    """torch.optim.Adam(betas=(0.9, 0.99)) semantics, one fused kernel per tensor (ffb_adam_step)."""

    def __init__(self, param_groups, betas=(0.9, 0.99), eps=1e-8):
        self.groups = []
        for g in param_groups:
            ps = [p for p in g['params']]
            self.groups.append({'params': ps, 'lr': g['lr']})
        self.betas, self.eps, self.t = betas, eps, 0
        self.state = {}
        for g in self.groups:
            for p in g['params']:
                self.state[p] = (torch.zeros_like(p), torch.zeros_like(p))

    @property
    def params(self):
        return [p for g in self.groups for p in g['params']]

    @torch.no_grad()
    def step(self, grads=None, grad_scale=1.0):
        """grads: list aligned with self.params (tensors sharing each parameter's strides) or None to use p.grad."""
        self.t += 1
        i = 0
        for g in self.groups:
            for p in g['params']:
                gr = grads[i] if grads is not None else p.grad
                i += 1
                if gr is None:
                    continue
                if gr.stride() != p.stride():
                    gr = gr.contiguous(memory_format=torch.channels_last_3d) if p.dim() == 5 else gr.contiguous()
                m, v = self.state[p]
                ops.adam_step(p, gr, m, v, g['lr'], self.betas[0], self.betas[1], self.eps, self.t, grad_scale)

    def decay_lr(self, factor):
        for g in self.groups:
            g['lr'] = g['lr'] * factor


class TrainStep:
    """One NeRF optimisation step of the reference's training loop (train_per_scene.py:149-171: render the ray batch,
    MSE, backward, Adam with two lr groups, multiplicative lr decay) as ONE replayable CUDA graph.

    B200-first design: every data-dependent size (valid samples, shaded samples) stays on the device
    (`model.lazy_counts`), so the whole step has a static launch sequence; it is captured once and replayed with a
    single `cudaGraphLaunch` per step — no Python, no allocator and no host synchronisation inside the step.  The
    optimiser scalars (step count, lr, bias corrections) live in device memory and are advanced by a kernel inside the
    graph.  Gradients accumulate directly into one flat fp32 arena (zeroed by one memset, all-reduced in place when
    world_size > 1, consumed by one multi-tensor Adam launch).

        ts = TrainStep(model, model.get_optparam_groups(lr_small, lr_large), batch=4096, n_samples=443, lr_decay=f)
        loss = ts.step(rays_host, rgb_host)          # host (pinned) or device tensors; returns a device scalar

    Schedule events that re-allocate factors or change the sample count (upsample / shrink / alpha-mask update) need a
    new TrainStep (`ts.rebuild()` keeps the optimiser scalars), exactly where the reference rebuilds its optimiser.
    """

    CHUNK = 4096

    def __init__(self, model, param_groups, batch, n_samples, white_bg=True, betas=(0.9, 0.99), eps=1e-8, lr_decay=1.0,
                 group=None, use_graph=True, warmup=2, nccl_in_graph=False, ndc_ray=False, overlap_comm=None, comm=None):
        """overlap_comm (opt-in): split the field backward in two phases and all-reduce the gradients of the first phase (fine
        basis levels + both MLPs) over NCCL while the second (coefficients + coarse levels) is still scattering.
        comm: 'symm' (default when world_size > 1: the arena lives in symmetric memory and is reduced by csrc/allreduce.cu inside
        the step's single CUDA graph) or 'nccl' (torch.distributed.all_reduce between two graphs; also the fallback when symmetric
        memory cannot be set up — the reason is printed)."""
        # measured on 2 and 8 B200s: the split costs more (+52 us of scatter, +3 launches) than the overlap hides -> opt-in only
        overlap_comm = bool(overlap_comm) and self._can_split(model)
        self.comm = (comm or os.environ.get('FFB_COMM', 'symm')).lower()
        self.model, self.B, self.S, self.white_bg = model, int(batch), int(n_samples), bool(white_bg)
        # NDC (llff) / unbounded (360) scenes: the interpx row shared by all rays is a static device buffer refreshed per step
        self.ndc_ray = bool(ndc_ray)
        self.z_kind = 'unbound' if getattr(model, 'is_unbound', False) else ('ndc' if self.ndc_ray else None)
        self.betas, self.eps, self.lr_decay, self.group, self.use_graph = betas, eps, float(lr_decay), group, use_graph
        self.nccl_in_graph = nccl_in_graph
        # frozen parameters (set_optimizable) stay out of the step, like torch.optim.Adam skips parameters without a gradient
        self.groups = [{'params': [p for p in g['params'] if p.requires_grad], 'lr': float(g['lr'])} for g in param_groups]
        self.params = [p for g in self.groups for p in g['params']]
        # Arena order (data parallel): the gradients that are complete LAST — the coefficient grid and the coarse basis levels,
        # scattered by the deferred second phase of the field backward — first, everything else behind them, so that both
        # all-reduce ranges are contiguous.  `late` = the tensors of that second phase.
        late_ids = {id(p) for p in (self._late_params(model) if overlap_comm else [])}
        self.late = [p for p in self.params if id(p) in late_ids]
        self.arena_order = [p for p in self.params if id(p) in late_ids] + [p for p in self.params if id(p) not in late_ids]
        dev = self.params[0].device
        self.dev = dev
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(group)
        # gradient arena + optimiser state
        self.symm = None
        if self.world > 1 and self.comm == 'symm' and not self.late:
            try:
                box = {}

                def alloc(n):
                    box['a'] = SymmArena(n, dev, group)
                    return box['a'].flat
                self.bucket = GradBucket(self.arena_order, flat_alloc=alloc)
                self.symm = box['a']
            except Exception as e:      # loud, not fatal: NCCL is an equally correct transport
                print(f'[ffb200] symmetric-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL', file=sys.stderr, flush=True)
                self.symm = None
        if self.symm is None:
            self.bucket = GradBucket(self.arena_order)
        self._arena_index = {id(p): k for k, p in enumerate(self.arena_order)}
        self.late_end = self.bucket.offsets[len(self.late)]          # arena[:late_end] = late gradients
        self.m = torch.zeros_like(self.bucket.flat)
        self.v = torch.zeros_like(self.bucket.flat)
        self.lr_d = torch.tensor([g['lr'] for g in self.groups], dtype=torch.float64, device=dev)
        self.step_d = torch.zeros(1, dtype=torch.int64, device=dev)
        self.hyper_d = torch.zeros(len(self.groups), 2, dtype=torch.float32, device=dev)
        self._build_tables()
        # static inputs / outputs
        self.rays_s = torch.zeros(self.B, 6, device=dev)
        self.target_s = torch.zeros(self.B, 3, device=dev)
        self.jitter_s = torch.zeros(self.B, device=dev)
        self.z_s = None
        if self.z_kind is not None:
            self.z_s = model._z_table_host(self.z_kind, self.S, False).to(dev)
        self.loss_s = torch.zeros(1, device=dev)
        # white-background decision as a device flag: for non-white-background scenes the reference flips a coin every step
        # (FactorFields.py:890); a captured graph must read it at replay time, not freeze the value seen at capture
        self.bg_s = torch.full((1,), int(self.white_bg), dtype=torch.int32, device=dev)
        self.graph = None
        self._warmup = warmup

    @staticmethod
    def _can_split(model):
        try:
            return bool(model._plan('coding').fast) and len(getattr(model, 'basises', [])) >= 2 and not isinstance(model.basises, torch.nn.ModuleList)
        except Exception:
            return False

    @staticmethod
    def _late_params(model):
        """Second phase of the field backward: the coefficient tensor(s) and the coarser half of the basis levels (by bytes
        the smaller part: 6 of 21 MB at nerf.yaml), so the bulk of the arena is on the wire while they scatter."""
        bas = list(model.basises)
        order = sorted(range(len(bas)), key=lambda i: bas[i].numel())
        total, acc, coarse = sum(b.numel() for b in bas), 0, []
        for i in order:
            if acc + bas[i].numel() > 0.35 * total:
                break
            coarse.append(bas[i])
            acc += bas[i].numel()
        return [p for p in list(model.coeffs) + coarse if p.requires_grad]

    def _build_tables(self):
        rows, ct, cs = [], [], []
        i = 0
        for gi, g in enumerate(self.groups):
            for p in g['params']:
                k = self._arena_index[id(p)]
                off, n = self.bucket.offsets[k], p.numel()
                base = off * 4
                rows.append([p.data_ptr(), self.bucket.flat.data_ptr() + base, self.m.data_ptr() + base, self.v.data_ptr() + base, n, gi])
                for s in range(0, n, self.CHUNK):
                    ct.append(i)
                    cs.append(s)
                i += 1
        self.table = torch.tensor(rows, dtype=torch.int64, device=self.dev)
        self.chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=self.dev)
        self.chunk_start = torch.tensor(cs, dtype=torch.int64, device=self.dev)
        self.arena = {p.data_ptr(): self.bucket.view(k) for k, p in enumerate(self.arena_order)}
        self._param_ptrs = [p.data_ptr() for p in self.params]

    # -- the step body: only stream-ordered device work (capturable) ------------------------------------------
    def _render_backward(self):
        """zero the arena, render, loss, backward: leaves this rank's gradients in self.bucket.flat"""
        from . import ops as _ops
        m = self.model
        self.bucket.flat.zero_()
        _ops.set_grad_arena(self.arena)
        prev_jitter = m.__dict__.get('_jitter')
        m._jitter = lambda n, tr: self.jitter_s
        m._z_static = self.z_s
        m._white_bg_static = self.bg_s
        prev_lazy = m.__dict__.get('lazy_counts', False)
        m.lazy_counts = True          # device-side sample counts, scoped to this step: direct model(rays) calls stay exact-sized
        self._split = {'late': {p.data_ptr() for p in self.late}} if self.late else None
        _ops.set_field_bwd_split(self._split)
        # in this graph linear_mat's output feeds the compositor and nothing else: its gradient can travel sparse (ops.SPARSE_FEAT_GRAD)
        prev_sparse = _ops.SPARSE_FEAT_GRAD
        _ops.SPARSE_FEAT_GRAD = self._sparse_ok()
        _ops.sparse_grads_pending()
        try:
            rgb, depth, _ = m(self.rays_s, white_bg=self.white_bg, is_train=True, ndc_ray=self.ndc_ray, N_samples=self.S)
            _, g_rgb = _ops.mse_fwd_bwd(rgb, self.target_s, loss=self.loss_s)
            grads = torch.autograd.grad([rgb], self.params, grad_outputs=[g_rgb], allow_unused=True)
            if _ops.sparse_grads_pending():
                raise RuntimeError('ffb200: a sparse compositor gradient was not consumed by linear_mat\'s backward')
        finally:
            _ops.SPARSE_FEAT_GRAD = prev_sparse
            _ops.set_grad_arena(None)
            _ops.set_field_bwd_split(None)
            m._z_static = None
            m._white_bg_static = None
            m.lazy_counts = prev_lazy
            if prev_jitter is None:
                m.__dict__.pop('_jitter', None)
            else:
                m._jitter = prev_jitter
        for k, g in enumerate(grads):       # gradients produced outside the arena (factor types without arena support)
            if g is not None:
                v = self.bucket.view(self._arena_index[id(self.params[k])])
                if g.data_ptr() != v.data_ptr():
                    v.copy_(g)

    def _sparse_ok(self):
        """True when linear_mat's backward can take the compositor's gradient in sparse form (ops.SPARSE_FEAT_GRAD): the plain
        2-layer MLPMixer of the per-scene configs in a shape the pipelined tensor-core kernel runs.  FFB_SPARSE_FEAT_GRAD=0 turns
        the hand-off off (dense gradient tensor, as a direct loss.backward() produces)."""
        if os.environ.get('FFB_SPARSE_FEAT_GRAD', '1') == '0':
            return False
        from . import native as nv
        lm = getattr(self.model, 'linear_mat', None)
        try:
            params, has_bias = lm._flat()
            if lm.pe != 0 or lm.with_dropout or len(has_bias) != 2 or tuple(has_bias) != (True, False):
                return False
            W1, W2 = params[0], params[2]
            if not all(p.requires_grad for p in params):
                return False
            return nv.lib().ffb_mlp2_pipelined_eligible(W1.shape[1], W1.shape[0], W2.shape[0]) == 1
        except Exception:
            return False

    def _late_backward(self):
        """Second phase of the field backward (deferred by _render_backward when `self.late` is non-empty)."""
        from . import ops as _ops
        if self._split:
            _ops.set_field_bwd_split(self._split)
            try:
                _ops.field_bwd_deferred()
            finally:
                _ops.set_field_bwd_split(None)

    def _all_reduce(self):
        if self.world > 1:
            if self.symm is not None:
                self.symm.all_reduce()
            else:
                torch.distributed.all_reduce(self.bucket.flat, op=torch.distributed.ReduceOp.SUM, group=self.group)

    def _all_reduce_early(self):
        """Asynchronous all-reduce of arena[late_end:] (complete after the first backward phase); -> work handle | None"""
        if self.world > 1 and self.late_end < self.bucket.flat.numel():
            return torch.distributed.all_reduce(self.bucket.flat[self.late_end:], op=torch.distributed.ReduceOp.SUM, group=self.group, async_op=True)
        return None

    def _all_reduce_late(self):
        if self.world > 1 and self.late_end > 0:
            torch.distributed.all_reduce(self.bucket.flat[:self.late_end], op=torch.distributed.ReduceOp.SUM, group=self.group)

    def _optimise(self):
        from . import ops as _ops
        _ops.adam_hyper_advance(self.lr_d, self.step_d, self.hyper_d, self.betas[0], self.betas[1], self.lr_decay)
        _ops.adam_multi(self.table, self.chunk_tensor, self.chunk_start, self.CHUNK, self.hyper_d, self.betas[0], self.betas[1],
                        self.eps, 1.0 / self.world)

    def _body(self):
        self._render_backward()
        if self.late:
            work = self._all_reduce_early()
            self._late_backward()
            self._all_reduce_late()
            if work is not None:
                work.wait()
        else:
            self._all_reduce()
        self._optimise()

    def _check_params(self):
        if [p.data_ptr() for p in self.params] != self._param_ptrs:
            raise RuntimeError('TrainStep: a parameter tensor was re-allocated (shrink / upsample); build a new TrainStep')

    def _capture(self):
        # warm-up on a side stream (allocator, lazy kernel attributes), then restore the state it touched
        snap = [p.detach().clone() for p in self.params]
        st = (self.m.clone(), self.v.clone(), self.lr_d.clone(), self.step_d.clone())
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(max(self._warmup, 1)):
                self._body()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        with torch.no_grad():
            for p, q in zip(self.params, snap):
                p.copy_(q)
            self.m.copy_(st[0]); self.v.copy_(st[1]); self.lr_d.copy_(st[2]); self.step_d.copy_(st[3])
        if not self.use_graph:
            self.graph = False
        elif self.world == 1 or self.nccl_in_graph or self.symm is not None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self.graph = (g,)
        else:
            # data parallel: [render + backward] graph -> NCCL all-reduce of the arena (a normal stream-ordered call
            # between the two replays) -> [Adam] graph
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga):
                self._render_backward()
            if self.late:
                # [render + backward, phase 1] -> async all-reduce of the early range || [backward phase 2] -> all-reduce of the
                # late range -> [Adam].  The stash of phase 2 keeps graph 1's buffers alive, so the shared pool cannot recycle them.
                gl = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gl, pool=ga.pool()):
                    self._late_backward()
                with torch.cuda.graph(gb, pool=ga.pool()):
                    self._optimise()
                self.graph = (ga, gl, gb)
            else:
                with torch.cuda.graph(gb, pool=ga.pool()):
                    self._optimise()
                self.graph = (ga, gb)

    def step(self, rays, target, jitter=None, bg_coin=None):
        """rays [B,6], target [B,3]: host (ideally pinned) or device fp32 tensors; jitter [B] (default: torch.rand on the
        CPU generator, one draw per ray, as FactorFields.py:593-595).  Returns the loss as a 1-element DEVICE tensor
        (valid until the next step); call .item() to read it back.  bg_coin: for white_bg=False scenes, the step's
        random-background decision (default: `torch.rand((1,)) < 0.5` on the CPU generator, FactorFields.py:890)."""
        if rays.shape[0] != self.B:
            raise RuntimeError(f'TrainStep was built for batches of {self.B} rays, got {rays.shape[0]}')
        self._check_params()
        if self.z_kind is not None:      # per-SAMPLE uniforms (FactorFields.py:579,612) instead of the per-ray jitter
            self.z_s.copy_(self.model._z_table_host(self.z_kind, self.S, True), non_blocking=True)
            jitter = self.jitter_s
        elif jitter is None:
            jitter = torch.rand(self.B, 1)[:, 0]
        self.rays_s.copy_(rays[:, :6], non_blocking=True)
        self.target_s.copy_(target, non_blocking=True)
        self.jitter_s.copy_(jitter, non_blocking=True)
        if not self.white_bg:       # the per-step coin of FactorFields.py:890, drawn after the jitter like the reference does
            self.bg_s.fill_(int(bool(torch.rand((1,)) < 0.5)) if bg_coin is None else int(bool(bg_coin)))
        if self.graph is None:      # capture warms up on THIS batch (an all-zero ray buffer would aim every scatter at one texel)
            self._capture()
        if not self.graph:
            self._body()
        elif len(self.graph) == 1:
            self.graph[0].replay()
        elif len(self.graph) == 3:
            self.graph[0].replay()
            work = self._all_reduce_early()
            self.graph[1].replay()
            self._all_reduce_late()
            if work is not None:
                work.wait()
            self.graph[2].replay()
        else:
            self.graph[0].replay()
            self._all_reduce()
            self.graph[1].replay()
        return self.loss_s

    def snapshot(self):
        """Copy of everything a step mutates (parameters, Adam moments, lr, step count)."""
        return ([p.detach().clone() for p in self.params], self.m.clone(), self.v.clone(), self.lr_d.clone(), self.step_d.clone())

    @torch.no_grad()
    def restore(self, snap):
        for p, q in zip(self.params, snap[0]):
            p.copy_(q)
        self.m.copy_(snap[1]); self.v.copy_(snap[2]); self.lr_d.copy_(snap[3]); self.step_d.copy_(snap[4])

    @property
    def lrs(self):
        return self.lr_d.tolist()

    def set_lrs(self, lrs):
        self.lr_d.copy_(torch.tensor(lrs, dtype=torch.float64))


class RegressStep(TrainStep):
    """One optimisation step of the reference's regression loops — 2-D image (scripts/2D_regression.ipynb cell 4), SDF
    (scripts/sdf_regression.ipynb cell 2) and image set (scripts/2D_set_regression.py:120-142):

        feats, _ = model.get_coding(x);  y = model.linear_mat(feats, is_train);  loss = mean((y - target)^2)
        (loss * loss_scale).backward();  Adam step;  loss_scale *= 0.1 ** (1 / n_iter)   [scale decays instead of the lr]

    as ONE replayable CUDA graph (field-query kernel -> fused MLP -> MSE -> MLP backward -> scatter -> multi-tensor
    Adam), with the loss scale kept in device memory.  x [B, d] are coordinates in aabb units (pixel centres /
    [0, 640]^3 points / (x, y, image+0.5)), target [B, out_dim].  `step` returns the UNSCALED loss like the notebooks print."""

    def __init__(self, model, param_groups, batch, x_dim, out_dim, loss_scale_decay=1.0, is_train=False, betas=(0.9, 0.99),
                 eps=1e-8, group=None, use_graph=True, warmup=2, local_params=()):
        """local_params: parameters that are NOT replicated across ranks (the image-set coefficient slabs each rank owns,
        `image_set_shard`): their gradients stay out of the all-reduce."""
        super().__init__(model, param_groups, batch, n_samples=1, betas=betas, eps=eps, lr_decay=1.0, group=group,
                         use_graph=use_graph, warmup=warmup, overlap_comm=False, comm='nccl')
        local = {p.data_ptr() for p in local_params}
        self._shared_ranges = arena_ranges(self.bucket, [p.data_ptr() not in local for p in self.arena_order])
        self.is_train, self.loss_scale_decay = bool(is_train), float(loss_scale_decay)
        self.rays_s = torch.zeros(self.B, int(x_dim), device=self.dev)         # coordinates
        self.target_s = torch.zeros(self.B, int(out_dim), device=self.dev)
        self.scale_d = torch.ones(1, dtype=torch.float64, device=self.dev)     # loss_scale (a Python double in the notebooks)
        self.scale_f = torch.ones(1, dtype=torch.float32, device=self.dev)

    def _render_backward(self):
        from . import ops as _ops
        m = self.model
        self.bucket.flat.zero_()
        _ops.scalar_decay(self.scale_d, self.loss_scale_decay, self.scale_f)    # `loss_scale *= lr_factor` opens the iteration
        _ops.set_grad_arena(self.arena)
        try:
            feats, _ = m.get_coding(self.rays_s)
            y = m.linear_mat(feats, is_train=self.is_train)
            _, g_y = _ops.mse_fwd_bwd(y.reshape(self.B, -1), self.target_s, loss=self.loss_s, g_scale_dev=self.scale_f)
            grads = torch.autograd.grad([y], self.params, grad_outputs=[g_y.view_as(y)], allow_unused=True)
        finally:
            _ops.set_grad_arena(None)
        for k, g in enumerate(grads):
            if g is not None:
                v = self.bucket.view(self._arena_index[id(self.params[k])])
                if g.data_ptr() != v.data_ptr():
                    v.copy_(g)

    def _all_reduce(self):
        if self.world > 1:
            for a, b in self._shared_ranges:     # replicated parameters only (contiguous arena ranges)
                torch.distributed.all_reduce(self.bucket.flat[a:b], op=torch.distributed.ReduceOp.SUM, group=self.group)

    def _capture(self):
        st = (self.scale_d.clone(), self.scale_f.clone())
        super()._capture()                      # the warm-up passes advanced the loss scale; captured replays start from `st`
        self.scale_d.copy_(st[0]); self.scale_f.copy_(st[1])

    def step(self, x, target):
        """x [B, d], target [B, out_dim]: host (ideally pinned) or device fp32 tensors -> loss (1-element device tensor)."""
        if x.shape[0] != self.B:
            raise RuntimeError(f'RegressStep was built for batches of {self.B} points, got {x.shape[0]}')
        self._check_params()
        self.rays_s.copy_(x, non_blocking=True)
        self.target_s.copy_(target.reshape(self.B, -1), non_blocking=True)
        if self.graph is None:
            self._capture()
        if not self.graph:
            self._body()
        elif len(self.graph) == 1:
            self.graph[0].replay()
        else:
            self.graph[0].replay()
            self._all_reduce()
            self.graph[1].replay()
        return self.loss_s

    def snapshot(self):
        return super().snapshot() + (self.scale_d.clone(), self.scale_f.clone())

    @torch.no_grad()
    def restore(self, snap):
        super().restore(snap[:5])
        self.scale_d.copy_(snap[5]); self.scale_f.copy_(snap[6])


@torch.no_grad()
def evaluate_field(model, coords, chunk=10240, is_train=False):
    """Dense evaluation of the regressed signal (eval_img of scripts/2D_regression.ipynb cell 1, eval_sdf /
    cal_l1_iou of scripts/sdf_regression.ipynb cell 1): model.linear_mat(model.get_coding(x)) over `coords` [N, d]
    (host or device) in chunks; returns a device tensor [N, out_dim]."""
    out = []
    for c in torch.split(coords, chunk, dim=0):
        feats, _ = model.get_coding(c.to(model.device, non_blocking=True).float())
        out.append(model.linear_mat(feats, is_train=is_train))
    return torch.cat(out)


def regression(cfg, model, coords, targets, n_iters=None, batch_size=None, index_fn=None, scale_loss=None, log=None,
               use_graph=True):
    """The regression drivers of the reference (scripts/2D_regression.ipynb, sdf_regression.ipynb,
    2D_set_regression.py) on a GPU-resident sample set: coords [N, d] / targets [N, out_dim] are moved to the device once
    (the notebooks' 8-worker DataLoader + per-step H2D copy is replaced by a device gather), every step draws
    `idx = torch.randint(0, N, (batch,))` on the CPU generator exactly like sdf_regression.ipynb cell 2 (or calls
    `index_fn(step)` -> LongTensor), and runs one RegressStep.  scale_loss: True for the image / sdf notebooks (loss *
    0.1**(k/n_iter)), False for the image set script (its scaling line is commented out); default by cfg.defaults.mode.
    Returns dict(loss=[per step], steps, seconds)."""
    import time
    t = cfg.training
    n_iters = int(n_iters if n_iters is not None else t.n_iters)
    B = int(batch_size if batch_size is not None else t.batch_size)
    if scale_loss is None:
        scale_loss = cfg.defaults.mode != 'images'
    dev = model.device
    coords_d = coords.to(dev).float().contiguous()
    targets_d = targets.to(dev).float().reshape(coords.shape[0], -1).contiguous()
    N = coords_d.shape[0]
    decay = 0.1 ** (1.0 / n_iters) if scale_loss else 1.0
    rs = RegressStep(model, model.get_optparam_groups(t.lr_small, t.lr_large), batch=B, x_dim=coords_d.shape[1],
                     out_dim=targets_d.shape[1], loss_scale_decay=decay, is_train=True, use_graph=use_graph)
    losses = []
    loss_hist = torch.zeros(n_iters, device=dev)
    torch.cuda.synchronize(dev)
    t0 = time.time()
    for it in range(n_iters):
        idx = index_fn(it) if index_fn is not None else torch.randint(0, N, (B,))
        idx = idx.to(dev, non_blocking=True)
        loss = rs.step(coords_d.index_select(0, idx), targets_d.index_select(0, idx))
        loss_hist[it:it + 1].copy_(loss)          # no host sync per step; the notebooks read the loss only to print it
        if log is not None and it % 100 == 0:
            log(f'Iteration {it:05d}: loss_dist = {float(loss.item()):.8f}')
    torch.cuda.synchronize(dev)
    seconds = time.time() - t0
    losses = loss_hist.tolist()
    return dict(loss=losses, steps=n_iters, seconds=seconds, step=rs)


# ----------------------------------------------------------------------------------------------------------------
# Training-loop host (train_per_scene.py:89-234 `reconstruction`): same schedule, same optimiser bookkeeping, same RNG
# consumption (numpy permutation sampler + one torch CPU uniform per ray), with the per-step work in TrainStep.
# Datasets, image writers and tensorboard are out of scope: rays / colours come in as tensors.
# ----------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def evaluate_psnr(model, rays, rgbs, white_bg=True, chunk=8192, N_samples=-1):
    """Mean-squared error -> PSNR of a forward-only render of `rays` (renderer.py:29-98 without the image writers)."""
    from .renderer import render_ray
    from .utils import mse2psnr
    lazy, model.lazy_counts = getattr(model, 'lazy_counts', False), False
    try:
        rgb_map, _ = render_ray(rays, model, chunk=chunk, N_samples=N_samples, white_bg=white_bg, is_train=False, device=model.device)
    finally:
        model.lazy_counts = lazy
    mse = float(torch.mean((rgb_map - rgbs.to(rgb_map.device)) ** 2))
    return mse2psnr(max(mse, 1e-12))


def reconstruction(cfg, model, allrays, allrgbs, white_bg=True, n_iters=None, test=None, log=None, use_graph=True):
    """Per-scene optimisation with the reference's schedule (train_per_scene.py:125-234).

    allrays [N,6] / allrgbs [N,3]: host tensors (the reference keeps them on the host, :146,151-152).  test: optional
    (rays, rgbs) evaluated at the end.  Returns dict(psnr_train=[...per step], psnr_test=float|None, steps=int).

    Under torch.distributed (one process per GPU) the loop is data parallel with weak scaling: every rank holds the same
    ray set and the same host random streams, each step draws ONE global batch of world * batch_size rays (+ their jitter)
    and rank r trains on slice r of it; TrainStep all-reduces the gradient arena, the alpha-mask lattice is evaluated
    sharded (getDenseAlpha), the final evaluation is sharded by ray.  The result equals a single process with batch size
    world * batch_size."""
    from .utils import N_to_reso, SimpleSampler, cal_n_samples, mse2psnr
    rank, world = _dist_world()
    t = cfg.training
    n_iters = int(n_iters if n_iters is not None else t.n_iters)
    decay_iters = t.lr_decay_iters if t.lr_decay_iters > 0 else t.n_iters
    lr_factor = t.lr_decay_target_ratio ** (1.0 / decay_iters)
    upsamp_list, mask_list, shrink_list = list(t.upsamp_list), list(t.update_AlphaMask_list), list(t.shrinking_list)
    reso_list = torch.linspace(t.volume_resoInit, t.volume_resoFinal, len(upsamp_list)).ceil().long().tolist()
    reso_cur = N_to_reso(t.volume_resoInit ** model.in_dim, model.aabb)
    n_samples = min(cfg.renderer.max_samples, cal_n_samples(reso_cur, cfg.renderer.step_ratio))
    global_batch = t.batch_size * world
    mine = shard_slice(global_batch, rank, world)
    sampler = SimpleSampler(allrays.shape[0], global_batch)
    pinned = allrays.is_pinned()
    ndc_ray = bool(getattr(cfg.dataset, 'ndc_ray', 0))

    def new_step(keep=None):
        ts = TrainStep(model, model.get_optparam_groups(t.lr_small, t.lr_large), batch=t.batch_size, n_samples=n_samples,
                       white_bg=white_bg, betas=(0.9, 0.99), lr_decay=lr_factor, use_graph=use_graph, ndc_ray=ndc_ray)
        if keep is not None:       # same parameters, new launch sequence (alpha mask changed): carry the optimiser over
            ts.m.copy_(keep.m); ts.v.copy_(keep.v); ts.lr_d.copy_(keep.lr_d); ts.step_d.copy_(keep.step_d)
        return ts

    ts = new_step()
    psnrs, reso_mask = [], None
    for it in range(n_iters):
        idx = sampler.nextids()[mine]
        rays_b, rgb_b = allrays[idx], allrgbs[idx]
        if pinned:
            rays_b, rgb_b = rays_b.pin_memory(), rgb_b.pin_memory()
        jitter = None if ts.z_kind is not None else torch.rand(global_batch, 1)[:, 0][mine]      # one uniform per ray (:593-595)
        loss_d = ts.step(rays_b, rgb_b, jitter)
        if world > 1:                                         # equal shards: the global MSE is the mean of the ranks' MSEs
            loss_d = loss_d.clone()
            torch.distributed.all_reduce(loss_d, op=torch.distributed.ReduceOp.SUM)
            loss_d = loss_d / world
        loss = float(loss_d.item())                           # the reference reads the loss every step too (:164)
        psnrs.append(mse2psnr(max(loss, 1e-12)))
        if log is not None and it % cfg.defaults.progress_refresh_rate == 0:
            log(f'Iteration {it:05d}: train_psnr = {psnrs[-1]:.2f} mse = {loss:.6f}')
        if it in mask_list or it in shrink_list:
            if reso_list and reso_list[0] < 256:
                reso_mask = N_to_reso(reso_list[0] ** model.in_dim, model.aabb)
            new_aabb = model.updateAlphaMask(tuple(reso_mask), is_update_alphaMask=it >= 1500)
            if it in shrink_list:
                model.shrink(new_aabb)
                ts = new_step()
            else:
                ts = new_step(keep=ts)
            if not cfg.dataset.ndc_ray and mask_list and it == mask_list[0] and not cfg.dataset.is_unbound:
                allrays, allrgbs = model.filtering_rays(allrays, allrgbs)
                sampler = SimpleSampler(allrgbs.shape[0], global_batch)
        if it in upsamp_list:
            n_voxels = reso_list.pop(0)
            reso_cur = N_to_reso(n_voxels ** model.in_dim, model.aabb)
            n_samples = min(cfg.renderer.max_samples, cal_n_samples(reso_cur, cfg.renderer.step_ratio))
            model.upsample_volume_grid(reso_cur)
            ts = new_step()
    out = dict(psnr_train=psnrs, psnr_test=None, steps=n_iters)
    if test is not None:
        if world > 1:
            lazy, model.lazy_counts = getattr(model, 'lazy_counts', False), False
            try:
                rgb_map, _ = render_sharded(test[0], model, chunk=8192, white_bg=white_bg, ndc_ray=ndc_ray)
            finally:
                model.lazy_counts = lazy
            out['psnr_test'] = mse2psnr(max(float(torch.mean((rgb_map - test[1].to(rgb_map.device)) ** 2)), 1e-12))
        else:
            out['psnr_test'] = evaluate_psnr(model, test[0], test[1], white_bg=white_bg)
    return out
