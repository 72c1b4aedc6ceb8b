#!/usr/bin/env python
"""bench.py — NeRF train rays/s (fwd + bwd + Adam) on the nerf.yaml shapes, synthetic data.

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels via libffb200.so)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (torch CPU operators, oracle/torch_port.py)

One JSON line on stdout (rank 0).  See the module docstring of bench_workload.py for the workload and DESIGN.md
§"Measurement" for how every number is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench_workload as W  # noqa: E402

METRIC = 'nerf_train_rays_per_s'
UNIT = 'rays/s'
WORKLOAD = ('nerf.yaml train step (fwd+bwd+Adam): 4096 rays x 443 samples/ray, aabb +-1, 5,347,600 params '
            '(coeff 18x32^3, basis {4,4,4,2,2,2}x{25..99}^3, linear_mat 18-64-32, renderModule 194-128-128-3), '
            'synthetic Blender-shaped rays, seeded synthetic mid-training weights (bench_workload.make_state)')
# algorithmic bytes per field query for nerf.yaml (fp32 storage)
B_FWD, B_BWD = 1236, 3528
B_BWD_OWN = 12 + 72 + 144 + 2 * 1152      # x + upstream gradient row + saved coeff/basis rows + read-modify-write scatter
FLOP_LINEAR_MAT, FLOP_RGB = 6400, 83200


class Variant:
    """What distinguishes the NeRF train-step workloads: nerf (the headline: nerf.yaml, grid x grid) and BASELINE config 4, the
    -vm / -CP presets on Tanks&Temples-shaped rays (non-cubic box, 1920x1080 rays, 498 samples per ray)."""

    def __init__(self, name):
        self.name = name
        if name == 'nerf':
            self.metric, self.workload, self.overrides = METRIC, WORKLOAD, {}
            self.aabb, self.n_samples, self.b_fwd, self.b_bwd, self.flop_lm = W.AABB, W.N_SAMPLES, B_FWD, B_BWD, FLOP_LINEAR_MAT
            self.rays = W.make_rays
        else:
            ov, fdim, bf, bb, fl = W.PRESETS[name]
            self.metric = name + '_train_rays_per_s'
            self.overrides, self.aabb, self.n_samples, self.b_fwd, self.b_bwd, self.flop_lm = ov, W.TNT_AABB, W.TNT_N_SAMPLES, bf, bb, fl
            self.rays = W.tnt_rays
            self.workload = (f'nerf.yaml + preset {name[5:]} ({", ".join(f"model.{k}={v}" for k, v in ov.items())}) train step (fwd+bwd+Adam): 4096 Tanks&Temples-shaped '
                             f'rays (1920x1080) x {W.TNT_N_SAMPLES} samples/ray, aabb {np.round(np.array(W.TNT_AABB), 2).tolist()}, Fdim {fdim}, seeded init + density offset '
                             '(bench_workload.density_offset_)')

    def cfg_overrides(self):
        return [f'model.{k}={json.dumps(v)}' for k, v in self.overrides.items()]

    def prepare(self, model):
        """Load / shape the synthetic state of a freshly built model (ours or the reference's)."""
        import torch
        if self.name == 'nerf':
            model.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state(0).items()})
            assert model.nSamples == 440 and float(model.stepSize) == float(W.step_size())
        else:
            W.density_offset_(model)
        return model


V = Variant('nerf')


def latest_field_capture():
    """The newest committed `ncu --set full` summary of the field kernels (profiles/rNN_ncu_field*.json) -> (dict, path)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_ncu_*.json')))
    for f in reversed(files):
        try:
            d = json.load(open(f))
            if any('fast_fwd' in k for k in d) and any('fast_bwd' in k for k in d):
                return d, os.path.relpath(f, ROOT)
        except Exception:
            pass
    return {}, None


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d['hbm_gbs']), tf=float(d.get('bf16_tflops_sustained', d['bf16_tflops'])), src='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tf=1400.0, src='fallback (B200_PROFILING.md)')


class ClockSampler:
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20', '-i', str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(', ') for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------------------
def _ref_stepper(device):
    """-> (step(rays_host, target_host, jitter_host), kind, stats()).  The UNMODIFIED reference (baseline/_ref/factor-fields, copied
    from /root/reference by __graft_entry__.build(); it travels to the GPU box with the working tree) through its own
    FactorFields + render_ray + torch.optim.Adam when the checkout is present, else the operator-level port
    (oracle/torch_port.py, kind "port")."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_step
    if ref_step.available():
        rs = ref_step.RefStep(None, V.aabb, device, V.n_samples, overrides=V.overrides)
        V.prepare(rs.model)

        def step(rays, target, jitter):
            # the reference draws its own per-ray jitter from the torch CPU generator (FactorFields.py:593-595)
            return rs.train_step(rays, target.to(device))
        return step, 'reference', lambda: rs.stats
    if V.name != 'nerf':
        raise RuntimeError('the operator-level port covers the grid x grid field only; the -vm / -CP presets need the reference checkout '
                           '(baseline/_ref/factor-fields, staged by __graft_entry__.build())')
    from oracle.torch_port import TorchPort
    tp = TorchPort(W.make_state(0), W.AABB, W.FREQ_BANDS, W.step_size(), W.RCFG, device=device)

    def step(rays, target, jitter):
        return tp.train_step(rays.to(device), target.to(device), W.N_SAMPLES, jitter.to(device))
    return step, 'port', lambda: tp.stats


REF_NOTE = ('ATen grid_sample parallelises over the batch axis, which the reference fixes at 1, so its dominant CPU op is effectively '
            'single-threaded whatever the core count')


def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores: the unmodified checkout when it travelled
    (kind "reference"), else the operator-level port.  Rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # stdout carries exactly ONE line (the JSON): the reference prints its parameter count to stdout, route fd 1 to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    torch.manual_seed(20211202)
    step, kind, stats = _ref_stepper('cpu')
    rays, target, jitter = V.rays(W.BATCH * 2, seed=1)
    rays, target, jitter = torch.from_numpy(rays), torch.from_numpy(target), torch.from_numpy(jitter)
    # bounded sample: size the per-step ray count so the whole run stays within a few minutes
    t0 = time.perf_counter()
    step(rays[:128], target[:128], jitter[:128])
    probe = time.perf_counter() - t0
    budget = 240.0
    n = int(min(W.BATCH, max(128, 128 * budget / max(probe, 1e-3) / max(args.steps + args.warmup, 1))))
    n = max(128, (n // 128) * 128)
    for i in range(args.warmup):
        s = (i * n) % (rays.shape[0] - n)
        step(rays[s:s + n], target[s:s + n], jitter[s:s + n])
    t0 = time.perf_counter()
    for i in range(args.steps):
        s = ((i + args.warmup) * n) % (rays.shape[0] - n)
        step(rays[s:s + n], target[s:s + n], jitter[s:s + n])
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = f'{n} of {W.BATCH} rays per step x {args.steps} steps (rays/s is per-ray, so the sample size does not bias it)'
    line = {'impl': 'reference', 'metric': V.metric, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': V.workload, 'rays_per_step': n},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample, 'note': REF_NOTE, 'stats': stats()},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + '\n').encode())


def comm_label(ts, world, eager):
    """config.parallelism: which transport all-reduces the gradient arena in this run."""
    mb = ts.bucket.flat.numel() * 4 / 1e6 if getattr(ts, 'bucket', None) is not None else 21.4
    symm = getattr(ts, 'symm', None)
    if symm is not None and not eager:
        path = 'multimem.ld_reduce / multimem.st through the NVSwitch' if getattr(symm, 'multicast', 0) else 'peer loads / stores over NVLink'
        return (f'ray-sharded dp{world}; the flat fp32 gradient arena ({mb:.1f} MB) all-reduced in place by one kernel over symmetric memory '
                f'({path}; csrc/allreduce.cu) inside the step\'s CUDA graph')
    return f'ray-sharded dp{world}, one NCCL all-reduce of the flat fp32 gradient arena ({mb:.1f} MB) per step'


def gpu_backlog(ms=60.0):
    """Park the stream behind a spinning kernel so that the Python-driven launches of the per-section profile pass queue up and then
    run back to back: the CUDA events around a section then bracket device time only (without it, sections of a few tens of
    microseconds — the regression drivers' kernels — mostly measured the host's launch gaps)."""
    import torch
    torch.cuda._sleep(int(ms * 1e-3 * 1.9e9))


def cpu_baseline_leg(n=256, steps=2):
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    step, kind, _ = _ref_stepper('cpu')
    rays, target, jitter = V.rays(n * (steps + 1), seed=1)
    rays, target, jitter = torch.from_numpy(rays), torch.from_numpy(target), torch.from_numpy(jitter)
    step(rays[:n], target[:n], jitter[:n])
    t0 = time.perf_counter()
    for i in range(1, steps + 1):
        step(rays[i * n:(i + 1) * n], target[i * n:(i + 1) * n], jitter[i * n:(i + 1) * n])
    dt = time.perf_counter() - t0
    what = 'the unmodified reference checkout, FactorFields + render_ray + torch.optim.Adam' if kind == 'reference' else 'oracle/torch_port.py'
    return {'value': n * steps / dt, 'unit': UNIT, 'cores': cores, 'kind': kind,
            'sample': f'{n} rays x {steps} steps of the same workload ({what}, torch CPU, {cores} threads)', 'note': REF_NOTE}


# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    # stdout carries exactly ONE line (the JSON): everything else that any library prints to fd 1 (NCCL's version banner,
    # the model's parameter count) is routed to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get('FFB_BENCH_WATCHDOG_S', '240')), exit=True)   # never hang a GPU box
    import torch
    import torch.distributed as dist
    import ffb200
    from ffb200 import native as nv
    from ffb200.models.FactorFields import FactorFields
    from ffb200.renderer import render_ray
    from ffb200.train import FusedAdam, GradBucket, TrainStep

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py (our arm) needs a CUDA device: there is no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')    # keep stdout for the one JSON line
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(20211202)
    np.random.seed(20211202)

    cfg = ffb200.load_cfg('nerf.yaml', V.cfg_overrides())
    cfg.dataset.aabb = V.aabb
    model = V.prepare(FactorFields(cfg, f'cuda:{local}'))
    model.lazy_counts = not args.exact_counts   # no host round trips for the data-dependent sample counts
    B, S = W.BATCH, V.n_samples
    if args.scaling == 'strong':       # fixed global batch: 4096 rays split across the ranks
        assert W.BATCH % world == 0
        B = W.BATCH // world

    groups = model.get_optparam_groups(cfg.training.lr_small, cfg.training.lr_large)
    lr_factor = 0.1 ** (1.0 / cfg.training.n_iters)
    eager = args.exact_counts or args.eager

    nb = 16                                                # distinct batches in the pool (weak scaling: 4096 rays per GPU)
    rays_np, target_np, jitter_np = V.rays(B * nb, seed=100 + rank)
    rays_h = torch.from_numpy(rays_np).pin_memory()
    target_h = torch.from_numpy(target_np).pin_memory()
    jitter_h = torch.from_numpy(jitter_np).pin_memory()
    rays_d, target_d, jitter_d = rays_h.to(dev), target_h.to(dev), jitter_h.to(dev)
    state = {'i': 0}

    if eager:
        # the pre-graph path: Python-driven launches, autograd + per-tensor fused Adam (kept for comparison)
        opt = FusedAdam(groups, betas=(0.9, 0.99))
        params = opt.params
        bucket = GradBucket(params) if world > 1 else None

        def optimise(loss):
            grads = torch.autograd.grad(loss, params, allow_unused=True)
            if bucket is not None:
                bucket.pack(grads)
                n = bucket.all_reduce()
                opt.step([bucket.view(i) for i in range(len(params))], grad_scale=1.0 / n)
            else:
                opt.step(list(grads))
            opt.decay_lr(lr_factor)

        def step_resident():
            b = state['i'] % nb
            state['i'] += 1
            sl = slice(b * B, (b + 1) * B)
            model._jitter = lambda n, tr: jitter_d[sl]
            rgb, depth, _ = model(rays_d[sl], white_bg=True, is_train=True, N_samples=S)
            loss = torch.mean((rgb - target_d[sl]) ** 2)
            optimise(loss)
            return loss

        def step_e2e():
            b = state['i'] % nb
            state['i'] += 1
            sl = slice(b * B, (b + 1) * B)
            model._jitter = lambda n, tr: jitter_h[sl].to(dev, non_blocking=True)
            rgb, depth, _ = render_ray(rays_h[sl], model, chunk=B, N_samples=S, white_bg=True, is_train=True, device=dev)
            loss = torch.mean((rgb - target_h[sl].to(dev, non_blocking=True)) ** 2)
            optimise(loss)
            return float(loss.item())

        step_profile = step_resident
        api = 'ffb200.renderer.render_ray(host rays) -> loss.item()'
    else:
        # the product path: the whole step (render, MSE, backward, all-reduce, Adam, lr decay) is one CUDA graph
        ov = os.environ.get('FFB_OVERLAP_COMM')
        ts = TrainStep(model, groups, batch=B, n_samples=S, white_bg=True, betas=(0.9, 0.99), lr_decay=lr_factor,
                       nccl_in_graph=bool(int(os.environ.get('FFB_NCCL_IN_GRAPH', '0'))), overlap_comm=None if ov is None else bool(int(ov)))

        def step_resident():
            """inputs already in HBM (device -> static-buffer copies only)"""
            b = state['i'] % nb
            state['i'] += 1
            sl = slice(b * B, (b + 1) * B)
            return ts.step(rays_d[sl], target_d[sl], jitter_d[sl])

        def step_e2e():
            """the user-facing call: host (pinned) rays / targets / jitter in, loss read back on the host"""
            b = state['i'] % nb
            state['i'] += 1
            sl = slice(b * B, (b + 1) * B)
            return float(ts.step(rays_h[sl], target_h[sl], jitter_h[sl]).item())

        def step_profile():
            """the same launch sequence outside the graph, so CUDA events can bracket each kernel group"""
            b = state['i'] % nb
            state['i'] += 1
            sl = slice(b * B, (b + 1) * B)
            ts.rays_s.copy_(rays_d[sl]); ts.target_s.copy_(target_d[sl]); ts.jitter_s.copy_(jitter_d[sl])
            ts._body()

        api = 'ffb200.train.TrainStep.step(host rays, host rgb, host jitter) -> loss.item()  [one CUDA-graph launch per step]'

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = nv.launch_count()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), nv.launch_count() - l0

    # Both timed regions run the SAME optimisation steps from the SAME model state (the per-step cost moves as the density
    # field trains: more samples pass the weight threshold), so `e2e` differs from `value` only by the host boundary.
    snap = None if eager else ts.snapshot()
    first = state['i']
    clocks = ClockSampler(local) if rank == 0 else None     # nvidia-smi needs ~0.1 s to start: launched before the warm-up (same load)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    ms_total, launches = timed(step_resident, args.steps)
    n_valid, n_app = int(model.last_stats['n_valid']), int(model.last_stats['n_app'])
    if snap is not None:
        ts.restore(snap)
        state['i'] = first
    for _ in range(max(args.warmup, 3)):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    clk = clocks.stop() if clocks else None

    # per-kernel device time of the same step (CUDA events on the launching stream), for the roofline
    step_profile()
    torch.cuda.synchronize()
    l0 = nv.launch_count()
    nv.profile_begin()
    P = 5
    gpu_backlog()
    for _ in range(P):
        step_profile()
    sec = {k: v[0] / P for k, v in nv.profile_end().items()}      # ms per step
    launches_per_step = (nv.launch_count() - l0) // P
    if not eager:
        launches = launches_per_step * args.steps                  # graph replays re-issue the captured launches

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = world * B * args.steps / (ms_total * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    kern = {}
    alg = {'field_fwd': n_valid * V.b_fwd, 'field_bwd': n_valid * V.b_bwd}
    for k, ms in sec.items():
        kern[k] = {'ms_per_step': round(ms, 4)}
        if k == 'field_fwd':
            kern[k]['achieved_GBps'] = round(alg[k] / (ms * 1e-3) / 1e9, 1)
            kern[k]['frac_hbm'] = round(alg[k] / (ms * 1e-3) / 1e9 / pk['hbm'], 4)
        elif k == 'field_bwd':       # no HBM fraction: the kernel is L2-reduction bound (see roofline_bwd)
            kern[k]['hbm_model_GBps'] = round(alg[k] / (ms * 1e-3) / 1e9, 1)
    flops = {'mlp_fwd': n_valid * V.flop_lm, 'mlp_bwd': 2 * n_valid * V.flop_lm, 'rgbmlp_fwd': n_app * FLOP_RGB,
             'rgbmlp_bwd': 2 * n_app * FLOP_RGB}
    for k, f in flops.items():
        if k in sec and sec[k] > 0:
            kern[k]['achieved_TFLOPs'] = round(f / (sec[k] * 1e-3) / 1e12, 3)
    # ---- roofline.  The HBM model applies to the GATHER kernel: every algorithmic byte of the forward query is a real
    # load or store.  The scatter kernel is bound by L2 reductions (ncu: DRAM ~12 %, red sectors the busiest unit) and is reported
    # against a live-measured L2-reduction throughput (roofline_bwd) instead of an HBM fraction it cannot be a fraction of.
    cap, cap_src = latest_field_capture() if V.name == 'nerf' else ({}, None)      # the committed captures are of the nerf.yaml kernels
    dom = 'field_fwd'
    ach = alg[dom] / (sec[dom] * 1e-3) / 1e9
    fwd_cap = next((rec for name, rec in cap.items() if 'fast_fwd' in name and rec.get('traffic_bytes')), None)
    roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': round(ach, 1), 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': round(ach / pk['hbm'], 4),
                'traffic': int(fwd_cap['traffic_bytes']) if fwd_cap else None,
                'traffic_source': (cap_src + ' (dram__bytes_read.sum + dram__bytes_write.sum, one launch)') if fwd_cap else None,
                'peak_source': pk['src'], 'algorithmic_bytes_per_query': V.b_fwd,
                'algorithmic_bytes_per_launch': alg[dom], 'queries_per_launch': n_valid, 'launch_ms': round(sec[dom], 4),
                'share_of_step': round(sec[dom] / (ms_total / args.steps), 4)}
    roofline_bwd = None
    try:
        sys.path.insert(0, os.path.join(ROOT, 'scratch'))
        import probe_red
        probe = probe_red.measure()
        bwd_cap = next((rec for name, rec in cap.items() if 'fast_bwd' in name and rec.get('lts__t_sectors_srcunit_tex_op_red.sum')), None)
        if bwd_cap and 'field_bwd' in sec:
            red = float(bwd_cap['lts__t_sectors_srcunit_tex_op_red.sum']['value'])
            q_cap = float(bwd_cap.get('queries_per_launch', 986959))
            sectors = red / q_cap * n_valid                 # the capture's red sectors per query x this launch's queries
            ach_b = sectors / (sec['field_bwd'] * 1e-3)
            peak_b = max(v['sectors_per_s'] for v in probe.values())
            roofline_bwd = {'kernel': 'field_bwd', 'bound': 'l2_red', 'unit': 'L2 reduction sectors/s', 'achieved': ach_b, 'peak': peak_b,
                            'frac': round(ach_b / peak_b, 4), 'red_sectors_per_query': round(red / q_cap, 2), 'sectors_source': cap_src +
                            ' (lts__t_sectors_srcunit_tex_op_red.sum, one launch)', 'peak_source': 'ffb_probe_red, measured in this run '
                            '(red.global.add.v4.f32 into an L2-resident 21 MB buffer; best of the two address patterns)', 'probe': probe,
                            'launch_ms': round(sec['field_bwd'], 4), 'share_of_step': round(sec['field_bwd'] / (ms_total / args.steps), 4),
                            'hbm_model_note': f'SURVEY 8(d) charges this kernel {V.b_bwd} B/query; it reads saved rows instead of re-gathering and its '
                                              'read-modify-write lands in L2, so an HBM fraction would be an accounting figure, not a bound'}
    except Exception as e:      # the probe is informational; the headline roofline never depends on it
        roofline_bwd = {'error': repr(e)}
    line = {'metric': V.metric, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': V.workload, 'rays_per_gpu_per_step': B, 'samples_per_ray': S, 'valid_fraction': round(n_valid / (B * S), 4),
                       'shaded_fraction_of_valid': round(n_app / max(n_valid, 1), 4), 'field_queries_per_step_per_gpu': n_valid,
                       'host_syncs_per_step': 2 if args.exact_counts else 0, 'cuda_graph': not eager, 'launches_per_step': launches_per_step,
                       'parallelism': (f'ray-sharded dp{world}; gradient arena all-reduced in two NCCL calls per step, the first (fine basis levels + MLPs, '
                                       f'{(ts.bucket.flat.numel() - ts.late_end) * 4 / 1e6:.1f} MB) overlapped with the second phase of the scatter, the second '
                                       f'({ts.late_end * 4 / 1e6:.1f} MB) exposed' if (not eager and ts.late) else comm_label(None if eager else ts, world, eager)) if world > 1 else 'single GPU',
                       'global_rays_per_step': world * B,
                       'l2': 'per-step inputs+intermediates (~0.5 GB) exceed the 126 MB L2; no explicit flush; the 21 MB of parameters stay '
                             'L2-resident across steps as in training'},
            'field_queries_per_s': world * n_valid * args.steps / (ms_total * 1e-3),
            'e2e': {'value': e2e, 'unit': UNIT, 'ms_per_step': ms_e2e / args.steps, 'h2d_bytes_per_step': B * (6 + 3 + 1) * 4,
                    'd2h_bytes_per_step': 4, 'api': api},
            'gpu_launches': launches, 'roofline': roofline, 'roofline_bwd': roofline_bwd, 'kernels': kern, 'clocks': clk}
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline_leg()
    if world == 1 and not args.no_cuda_eager_baseline:
        line['reference_cuda_eager'] = cuda_eager_leg()
        line['vs_reference_cuda_eager'] = round(e2e / line['reference_cuda_eager']['value'], 2)
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + '\n').encode())
    faulthandler.cancel_dump_traceback_later()
    if world > 1:
        dist.destroy_process_group()


def cuda_eager_leg(steps=5):
    """the reference's eager CUDA path on the same GPU is the kernel-level bar.  Runs the unmodified reference
    (else the operator-level port) with device='cuda' on the full 4096-ray batch: host rays in (render_ray copies them, as in
    train_per_scene.py:151-156), loss read back every step (:164)."""
    import torch
    step, kind, stats = _ref_stepper('cuda')
    rays, target, jitter = V.rays(W.BATCH * (steps + 2), seed=1)
    rays, target, jitter = (torch.from_numpy(a) for a in (rays, target, jitter))
    sl = lambda i: slice(i * W.BATCH, (i + 1) * W.BATCH)
    for i in range(2):
        step(rays[sl(i)], target[sl(i)], jitter[sl(i)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(2, steps + 2):
        step(rays[sl(i)], target[sl(i)], jitter[sl(i)])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {'value': W.BATCH / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms,
            'kind': kind + ' on CUDA (torch eager operators, fp32, TF32 off as in the reference)',
            'sample': f'{W.BATCH} rays x {steps} steps, host rays in, loss read back every step like train_per_scene.py:164',
            'stats': stats()}


# ------------------------------------------------------------------------------------------------------------
# Regression workloads (BASELINE.md §1 holds the notebooks' recorded rates for image and sdf)
# ------------------------------------------------------------------------------------------------------------
REGRESS = {
    # name: (yaml, aabb, batch, x_dim, out_dim, B_fwd, B_bwd, MLP flop/query, published queries/s, source)
    'image': ('image.yaml', [[0., 0.], [1024., 1024.]], 102400, 2, 3, 1736, 4032, 18816, 11.1e6, 'scripts/2D_regression.ipynb:140 (108.54 it/s x 102400, unstated GPU, incl. DataLoader)'),
    'sdf': ('sdf.yaml', [[0., 0., 0.], [640., 640., 640.]], 40960, 3, 1, 1236, 3528, 36, 13.3e6, 'scripts/sdf_regression.ipynb:215-302 (316.7-333.2 it/s x 40960, unstated GPU)'),
    'image_set': ('image_set.yaml', [[0, 0, 0], [256, 256, 800]], 40960, 3, 3, 5196, 14400, 18816, None, None),
}


def regress_inputs(name, n, seed):
    """Synthetic seeded sample points of the workload's shape: pixel centres / uniform points / (pixel centre, image+0.5)."""
    yaml_, aabb, B, xd, od = REGRESS[name][:5]
    rng = np.random.RandomState(seed)
    hi = np.array(aabb[1], np.float64)
    if name == 'sdf':
        x = rng.uniform(0, hi, (n, 3))
    else:
        x = np.floor(rng.uniform(0, hi, (n, xd))) + 0.5
    t = rng.uniform(0, 1, (n, od)) if name != 'sdf' else rng.uniform(-0.5, 0.5, (n, od))
    return x.astype(np.float32), t.astype(np.float32)


def regress_cpu_leg(name, state, model_facts, n, steps):
    import torch
    from oracle.torch_port import RegressPort
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    yaml_ = REGRESS[name][0]
    mode = 'nearest' if name == 'image' else 'bilinear'
    rp = RegressPort(state, model_facts['aabb'], model_facts['freq_bands'], model_facts['in_dim'], mode, mode, 0.002, 0.002, 1.0)
    x, t = regress_inputs(name, n * (steps + 1), 7)
    x, t = torch.from_numpy(x), torch.from_numpy(t)
    rp.train_step(x[:n], t[:n])
    t0 = time.perf_counter()
    for i in range(1, steps + 1):
        rp.train_step(x[i * n:(i + 1) * n], t[i * n:(i + 1) * n])
    dt = time.perf_counter() - t0
    return n * steps / dt, cores, f'{n} points x {steps} steps of the same workload (oracle/torch_port.RegressPort, torch CPU, {cores} threads)'


def run_regress(args):
    """`--workload image|sdf|image_set`: one regression train step (get_coding -> linear_mat -> MSE -> backward -> Adam) per
    step; metric = field queries/s.  Not the headline (BASELINE.json's metric is the nerf.yaml step): an extra line for the
    2-D / SDF / image-set drivers, whose notebooks hold the reference's only recorded rates."""
    name = args.workload
    yaml_, aabb, B, xd, od, b_fwd, b_bwd, mlp_flop, published, pub_src = REGRESS[name]
    rank = int(os.environ.get('RANK', '0'))
    metric, unit = f'{name}_regression_field_queries_per_s', 'queries/s'
    if args.impl == 'reference':
        if rank != 0:
            return
        import torch
        import ffb200
        from ffb200.models.FactorFields import field_shapes
        cfg = ffb200.load_cfg(yaml_)
        if name == 'image_set':
            aabb = [[0, 0, 0], [256, 256, 16]]      # 16 of the 800 coefficient slabs: per-query cost does not depend on the slab count
        sh = field_shapes(cfg, aabb)
        state = W.regress_state(cfg, sh, seed=0)
        facts = dict(aabb=sh['aabb'].numpy(), freq_bands=sh['freq_bands'].numpy(), in_dim=sh['in_dim'])
        n = 8192
        val, cores, sample = regress_cpu_leg(name, state, facts, n, max(args.steps, 2))
        line = {'impl': 'reference', 'metric': metric, 'value': val, 'unit': unit, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': n / val * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': f'{yaml_} regression step', 'points_per_step': n},
                'cpu_baseline': {'value': val, 'unit': unit, 'cores': cores, 'kind': 'port', 'sample': sample},
                'e2e': {'value': val, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
        print(json.dumps(line), flush=True)
        return
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get('FFB_BENCH_WATCHDOG_S', '300')), exit=True)
    import torch
    import torch.distributed as dist
    import ffb200
    from ffb200 import native as nv
    from ffb200.models.FactorFields import FactorFields
    from ffb200.train import RegressStep
    world, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(20211202)
    cfg = ffb200.load_cfg(yaml_, ['model.with_dropout=false'] if args.no_dropout else [])
    cfg.dataset.aabb = aabb
    model = FactorFields(cfg, f'cuda:{local}')
    with torch.no_grad():      # mid-training-like factors (coef_init is a constant; the kernels' cost does not depend on the values)
        for p in model.coeffs:
            p.add_(0.05 * torch.randn_like(p))
    nb = 8
    x_np, t_np = regress_inputs(name, B * nb, 100 + rank)
    x_h, t_h = torch.from_numpy(x_np).pin_memory(), torch.from_numpy(t_np).pin_memory()
    x_d, t_d = x_h.to(dev), t_h.to(dev)
    rs = RegressStep(model, model.get_optparam_groups(cfg.training.lr_small, cfg.training.lr_large), batch=B, x_dim=xd, out_dim=od,
                     loss_scale_decay=0.1 ** (1.0 / cfg.training.n_iters) if name != 'image_set' else 1.0, is_train=True)
    state = {'i': 0}

    def sl():
        b = state['i'] % nb
        state['i'] += 1
        return slice(b * B, (b + 1) * B)

    def step_resident():
        s = sl()
        return rs.step(x_d[s], t_d[s])

    def step_e2e():
        s = sl()
        return float(rs.step(x_h[s], t_h[s]).item())

    def step_profile():
        s = sl()
        rs.rays_s.copy_(x_d[s]); rs.target_s.copy_(t_d[s])
        rs._body()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    snap = None
    for _ in range(max(args.warmup, 3)):
        step_resident()
    snap = rs.snapshot()
    clocks = ClockSampler(local) if rank == 0 else None
    ms_total = timed(step_resident, args.steps)
    rs.restore(snap)
    for _ in range(max(args.warmup, 3)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clk = clocks.stop() if clocks else None
    step_profile()
    torch.cuda.synchronize()
    l0 = nv.launch_count()
    nv.profile_begin()
    P = 5
    gpu_backlog()
    for _ in range(P):
        step_profile()
    sec = {k: v[0] / P for k, v in nv.profile_end().items()}
    launches_per_step = (nv.launch_count() - l0) // P
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = world * B * args.steps / (ms_total * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    alg = {'field_fwd': B * b_fwd, 'field_bwd': B * b_bwd}
    kern = {}
    for k, ms in sec.items():
        kern[k] = {'ms_per_step': round(ms, 4)}
        if k in alg:
            kern[k]['achieved_GBps'] = round(alg[k] / (ms * 1e-3) / 1e9, 1)
            kern[k]['frac_hbm'] = round(alg[k] / (ms * 1e-3) / 1e9 / pk['hbm'], 4)
    dom = max(('field_fwd', 'field_bwd'), key=lambda k: sec.get(k, 0.0))
    ach = alg[dom] / (sec[dom] * 1e-3) / 1e9
    n_params = sum(p.numel() for p in model.parameters())
    line = {'metric': metric, 'value': value, 'unit': unit, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': (value / published) if published else None, 'baseline_source': pub_src, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{yaml_} regression train step (get_coding -> linear_mat -> MSE -> backward -> Adam), batch {B} points, '
                                   f'aabb {aabb}, {n_params:,} params, synthetic seeded points/targets',
                       'points_per_gpu_per_step': B, 'cuda_graph': True, 'launches_per_step': launches_per_step,
                       'l2': 'the factors are L2-resident (as in training); per-step activations stream through HBM; no explicit flush'},
            'e2e': {'value': e2e, 'unit': unit, 'ms_per_step': ms_e2e / args.steps, 'h2d_bytes_per_step': B * (xd + od) * 4, 'd2h_bytes_per_step': 4,
                    'api': 'ffb200.train.RegressStep.step(host x, host target) -> loss.item()  [one CUDA-graph launch per step]'},
            'gpu_launches': launches_per_step * args.steps,
            'roofline': {'kernel': dom, 'bound': 'hbm', 'achieved': round(ach, 1), 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': round(ach / pk['hbm'], 4),
                         'traffic': None, 'peak_source': pk['src'], 'algorithmic_bytes_per_launch': alg[dom], 'queries_per_launch': B,
                         'launch_ms': round(sec[dom], 4), 'share_of_step': round(sec[dom] / (ms_total / args.steps), 4)},
            'kernels': kern, 'clocks': clk}
    if world == 1 and not args.no_cpu_baseline:
        from ffb200.models.FactorFields import field_shapes
        cpu_aabb = [[0, 0, 0], [256, 256, 16]] if name == 'image_set' else aabb
        sh = field_shapes(cfg, cpu_aabb)
        v, cores, sample = regress_cpu_leg(name, W.regress_state(cfg, sh, seed=0),
                                           dict(aabb=sh['aabb'].numpy(), freq_bands=sh['freq_bands'].numpy(), in_dim=sh['in_dim']), 8192, 2)
        line['cpu_baseline'] = {'value': v, 'unit': unit, 'cores': cores, 'kind': 'port', 'sample': sample}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + '\n').encode())
    faulthandler.cancel_dump_traceback_later()
    if world > 1:
        dist.destroy_process_group()


def run_eval(args):
    """`--workload nerf_eval`: forward-only rendering of 800x800 test images (renderer.py:29-98 `evaluation` without the image
    writers): one step = one image = 640 000 rays through ffb200.renderer.render_ray, eval sampling (nSamples = 440, no jitter)."""
    rank = int(os.environ.get('RANK', '0'))
    metric, unit, H = 'nerf_eval_rays_per_s', 'rays/s', 800
    n_rays = H * H
    if args.impl == 'reference':
        if rank != 0:
            return
        import torch
        from oracle.torch_port import TorchPort
        cores = os.cpu_count()
        torch.set_num_threads(cores)
        tp = TorchPort(W.make_state(0), W.AABB, W.FREQ_BANDS, W.step_size(), W.RCFG)
        rays = torch.from_numpy(W.make_rays(1024 * (args.steps + 1), seed=1)[0])
        with torch.no_grad():
            tp.forward(rays[:1024], 440, None)
            t0 = time.perf_counter()
            for i in range(1, args.steps + 1):
                tp.forward(rays[i * 1024:(i + 1) * 1024], 440, None)
            dt = time.perf_counter() - t0
        val = 1024 * args.steps / dt
        line = {'impl': 'reference', 'metric': metric, 'value': val, 'unit': unit, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': {'workload': 'nerf.yaml forward-only render', 'rays_per_step': 1024},
                'cpu_baseline': {'value': val, 'unit': unit, 'cores': cores, 'kind': 'port',
                                 'sample': f'1024-ray chunks (the reference\'s evaluation chunk, renderer.py:50) x {args.steps}'},
                'e2e': {'value': val, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
        print(json.dumps(line), flush=True)
        return
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import ffb200
    from ffb200 import native as nv
    from ffb200.models.FactorFields import FactorFields
    from ffb200.renderer import render_ray
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    cfg = ffb200.load_cfg('nerf.yaml')
    cfg.dataset.aabb = W.AABB
    model = FactorFields(cfg, f'cuda:{local}')
    model.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state(0).items()})
    chunk = args.eval_chunk
    imgs = [torch.from_numpy(W.make_rays(n_rays, seed=200 + i)[0]).pin_memory() for i in range(2)]
    imgs_d = [r.to(dev) for r in imgs]
    state = {'i': 0}

    def step_resident():
        state['i'] += 1
        with torch.no_grad():
            return render_ray(imgs_d[state['i'] % 2], model, chunk=chunk, N_samples=-1, white_bg=True, is_train=False, device=dev)[0]

    def step_e2e():
        state['i'] += 1
        with torch.no_grad():
            rgb, depth = render_ray(imgs[state['i'] % 2], model, chunk=chunk, N_samples=-1, white_bg=True, is_train=False, device=dev)
        return rgb.cpu(), depth.cpu()                      # the reference moves both maps to the host per image (renderer.py:54)

    def timed(fn, k):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = nv.launch_count()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), nv.launch_count() - l0

    for _ in range(max(args.warmup, 3)):
        step_resident()
    clocks = ClockSampler(local)
    ms, launches = timed(step_resident, args.steps)
    for _ in range(3):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    clk = clocks.stop()
    # per-kernel device time over the chunks of ONE image (render_ray's loop restated so that the image's field queries can be counted)
    nv.profile_begin()
    n_valid = 0
    with torch.no_grad():
        img = imgs_d[0]
        for start in range(0, n_rays, chunk):
            model(img[start:start + chunk], is_train=False, white_bg=True, ndc_ray=False, N_samples=-1)
            n_valid += int(model.last_stats['n_valid'])
    sec = {k: v[0] for k, v in nv.profile_end().items()}
    pk = peaks()
    value, e2e = n_rays * args.steps / (ms * 1e-3), n_rays * args.steps / (ms_e2e * 1e-3)
    line = {'metric': metric, 'value': value, 'unit': unit, 'n_gpus': 1, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'nerf.yaml forward-only render of 800x800 images ({n_rays} rays x {model.nSamples} samples, no jitter), '
                                   f'render_ray chunks of {chunk} rays (the reference uses 1024 / 8192, renderer.py:50,124), bench_workload.make_state weights',
                       'chunk': chunk, 'l2': 'each image streams ~6 GB of samples / features through HBM; parameters stay L2-resident'},
            'e2e': {'value': e2e, 'unit': unit, 'ms_per_step': ms_e2e / args.steps, 'h2d_bytes_per_step': n_rays * 24, 'd2h_bytes_per_step': n_rays * 16,
                    'api': 'ffb200.renderer.render_ray(host rays) -> rgb_map.cpu(), depth_map.cpu()'},
            'gpu_launches': launches,
            'roofline': {'kernel': 'field_fwd', 'bound': 'hbm', 'achieved': round(n_valid * B_FWD / (sec['field_fwd'] * 1e-3) / 1e9, 1), 'peak': pk['hbm'],
                         'unit': 'GB/s', 'frac': round(n_valid * B_FWD / (sec['field_fwd'] * 1e-3) / 1e9 / pk['hbm'], 4), 'traffic': None,
                         'peak_source': pk['src'], 'note': 'field forward summed over the chunks of one image.  SURVEY 8(d) charges the 1 080 B of texel gathers per query to HBM; the parameters are L2-resident, so the kernel really streams x + two rows (156 B per query) and this fraction can exceed 1', 'queries_per_image': n_valid,
                         'launch_ms': round(sec['field_fwd'], 4), 'share_of_step': round(sec['field_fwd'] / (ms / args.steps), 4)},
            'kernels': {k: {'ms_per_step': round(v, 4)} for k, v in sec.items()}, 'clocks': clk}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + '\n').encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--eval-chunk', type=int, default=65536)
    ap.add_argument('--workload', default='nerf', choices=['nerf', 'nerf_vm', 'nerf_cp', 'nerf_eval'] + list(REGRESS), help='nerf (the headline, default) or a regression driver')
    ap.add_argument('--no-dropout', action='store_true', help='image_set: disable F.dropout on the MLP input')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'], help='weak: 4096 rays per GPU (default); strong: 4096 rays in total')
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-cuda-eager-baseline', action='store_true', help='skip timing the reference (eager torch CUDA operators) on the same GPU')
    ap.add_argument('--eager', action='store_true', help='Python-driven launches instead of the CUDA-graph TrainStep')
    ap.add_argument('--exact-counts', action='store_true', help='read the sample counts back every step (the reference-like sync mode)')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'ours' and args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called as plain `python bench.py --gpus N`
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}', '--master-addr', '127.0.0.1',
               '--master-port', str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    global V
    if args.workload in ('nerf_vm', 'nerf_cp'):
        V = Variant(args.workload)
    if args.workload == 'nerf_eval':
        run_eval(args)
    elif args.workload in REGRESS:
        run_regress(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
