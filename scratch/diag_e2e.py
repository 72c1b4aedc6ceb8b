import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workload as W
import ffb200
from ffb200 import native as nv
from ffb200.models.FactorFields import FactorFields
from ffb200.renderer import render_ray
from ffb200.train import FusedAdam
dev = torch.device('cuda', 0)
cfg = ffb200.load_cfg('nerf.yaml'); cfg.dataset.aabb = W.AABB
model = FactorFields(cfg, 'cuda:0')
model.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state(0).items()})
lazy = len(sys.argv) < 2 or sys.argv[1] != 'exact'
model.lazy_counts = lazy
B, S = W.BATCH, W.N_SAMPLES
opt = FusedAdam(model.get_optparam_groups(0.001, 0.02)); params = opt.params
nb = 16
rays_np, target_np, jitter_np = W.make_rays(B * nb, seed=100)
rays_h = torch.from_numpy(rays_np).pin_memory(); target_h = torch.from_numpy(target_np).pin_memory(); jitter_h = torch.from_numpy(jitter_np).pin_memory()
rays_d, target_d, jitter_d = rays_h.to(dev), target_h.to(dev), jitter_h.to(dev)
st = {'i': 0}
def T(): return time.perf_counter()
def step(e2e, sync, marks=None):
    b = st['i'] % nb; st['i'] += 1; sl = slice(b * B, (b + 1) * B)
    t0 = T()
    if e2e:
        model._jitter = lambda n, tr: jitter_h[sl].to(dev, non_blocking=True)
        rgb, depth, _ = render_ray(rays_h[sl], model, chunk=B, N_samples=S, white_bg=True, is_train=True, device=dev)
        tgt = target_h[sl].to(dev, non_blocking=True)
    else:
        model._jitter = lambda n, tr: jitter_d[sl]
        rgb, depth, _ = model(rays_d[sl], white_bg=True, is_train=True, N_samples=S)
        tgt = target_d[sl]
    t1 = T()
    loss = torch.mean((rgb - tgt) ** 2)
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    t2 = T()
    opt.step(list(grads)); opt.decay_lr(0.999)
    t3 = T()
    if sync: v = float(loss.item())
    t4 = T()
    if marks is not None: marks.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
for mode in [(False, False), (False, True), (True, False), (True, True)]:
    for _ in range(5): step(*mode)
    torch.cuda.synchronize()
    m = []
    t0 = T()
    for _ in range(20): step(*mode, marks=m)
    torch.cuda.synchronize()
    dt = (T() - t0) / 20 * 1e3
    a = np.array(m).mean(0) * 1e3
    ms = torch.cuda.memory_stats()
    print(f'lazy={lazy} e2e={mode[0]} sync={mode[1]}: {dt:.2f} ms/step wall; cpu fwd {a[0]:.2f} bwd {a[1]:.2f} adam {a[2]:.2f} item {a[3]:.2f}; '
          f'cudaMalloc segs {ms["num_device_alloc"]} frees {ms["num_device_free"]} retries {ms["num_alloc_retries"]} reserved {ms["reserved_bytes.all.peak"]/2**30:.1f} GiB', flush=True)
if lazy:
    from torch.profiler import profile, ProfilerActivity
    for _ in range(3): step(False, True)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3): step(False, True)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    # print the kernel timeline of the last step: name, start offset, duration
    t_end = ev[-1].time_range.end
    last = [e for e in ev if e.time_range.start > t_end - 36000]
    t0 = last[0].time_range.start
    for e in last:
        print(f'{(e.time_range.start - t0)/1e3:9.3f} ms  {e.time_range.elapsed_us():9.1f} us  {e.name[:90]}')
