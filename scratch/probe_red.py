"""Measured L2 reduction throughput (ffb_probe_red): 16-byte vector reductions per second into an L2-resident buffer of the
gradient arena's size.  Prints one JSON object."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ffb200 import native as nv


def measure(n_floats=5347712, blocks=148 * 8, iters=512, reps=5):
    buf = torch.zeros(n_floats, device='cuda')
    out = {}
    for pattern, name in ((0, 'random_16B_slots'), (1, 'warp_contiguous_512B')):
        n_ops = C.c_int64()
        best = None
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nv.check(nv.lib().ffb_probe_red(nv.ptr(buf), C.c_int64(n_floats), blocks, iters, pattern, C.byref(n_ops), nv.stream()))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        ops_s = n_ops.value / (best * 1e-3)
        sectors = 1.0 if pattern == 0 else 0.5          # 32-byte L2 sectors touched per 16-byte reduction
        out[name] = {'ms': round(best, 4), 'vector_reductions_per_s': ops_s, 'GB_per_s': ops_s * 16 / 1e9, 'sectors_per_s': ops_s * sectors}
    return out


if __name__ == '__main__':
    print(json.dumps(measure()))
