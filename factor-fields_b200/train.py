"""Train-step glue (train_per_scene.py:124-171): fused Adam over the model's parameter groups with the reference's
per-step multiplicative lr decay, and the flat gradient bucket used for the data-parallel all-reduce."""
import torch

from . import ops


def shard_slice(n, rank, world):
    """Contiguous slice of a global batch of n rays (or sample points) owned by `rank`: disjoint, ordered, covering
    [0, n); the first n % world ranks get one extra element (every rank draws the same global permutation
    and takes its own slice)."""
    base, rem = divmod(int(n), int(world))
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


class GradBucket:
    """Flat fp32 buffer holding every gradient back to back (parameter storage order), so that one all-reduce per
    step covers grids + MLPs.  Device-agnostic torch code (tested with gloo on CPU)."""

    def __init__(self, params):
        self.params = [p for p in params]
        self.sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for n in self.sizes:
            self.offsets.append(self.offsets[-1] + n)
        dev = self.params[0].device
        self.flat = torch.zeros(self.offsets[-1], device=dev, dtype=torch.float32)

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
