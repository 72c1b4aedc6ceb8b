"""N-GPU check + timing of the symmetric-memory all-reduce kernel (csrc/allreduce.cu) against NCCL on the nerf.yaml arena size:
torchrun --nproc-per-node N tests/dist/dist_check_allreduce.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from ffb200.train import SymmArena
local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
n = 5347712
ok = True
for nvls in ('1', '0'):
    os.environ['FFB_ALLREDUCE_NVLS'] = nvls
    try:
        ar = SymmArena(n, dev)
    except Exception as e:
        print(f'rank {rank}: SymmArena failed: {type(e).__name__}: {e}', flush=True)
        ok = False
        break
    path = 'NVLS multimem' if ar.multicast else 'P2P loads/stores'
    for trial in range(3):
        g = torch.Generator(device=dev).manual_seed(100 * trial + rank)
        x = torch.randn(n, device=dev, generator=g)
        ar.flat.copy_(x)
        ref = x.clone()
        dist.all_reduce(ref)
        torch.cuda.synchronize(); dist.barrier()
        ar.all_reduce()
        torch.cuda.synchronize()
        err = float((ar.flat - ref).abs().max() / ref.abs().max())
        # every rank must hold bit-identical sums
        chk = ar.flat.double().sum().reshape(1).clone()
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        same = all(float(v) == float(lst[0]) for v in lst)
        ok = ok and err < 1e-6 and same
        if rank == 0:
            print(f'{path}: trial {trial}: max rel err vs NCCL {err:.2e}, identical on all ranks: {same}', flush=True)
    def timeit(fn, reps=50):
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    t_ours = timeit(ar.all_reduce)
    for nb in (32, 64, 148):
        ar.BLOCKS = nb
        tt = timeit(ar.all_reduce)
        if rank == 0:
            print(f'   blocks {nb}: {tt:.1f} us', flush=True)
    buf = torch.randn(n, device=dev)
    t_nccl = timeit(lambda: dist.all_reduce(buf))
    if rank == 0:
        print(f'{path}: {n * 4 / 1e6:.1f} MB fp32 over {world} GPUs: ours {t_ours:.1f} us, NCCL {t_nccl:.1f} us', flush=True)
    if not ar.multicast and nvls == '1':
        break           # no multicast on this system: the second pass would repeat the P2P path
print(f'rank {rank}: {"OK" if ok else "MISMATCH"}', flush=True)
dist.barrier()
dist.destroy_process_group()
assert ok
