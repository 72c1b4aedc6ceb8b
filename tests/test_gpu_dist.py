"""Multi-GPU (NCCL) checks of the data-parallel paths, one process per GPU.  Skipped on boxes with one GPU;
the host-side logic of the same paths is covered on CPU by tests/test_dist_gloo.py (gloo, world size 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, nproc=2, timeout=600, args=()):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nproc}', '--master-addr', '127.0.0.1',
           '--master-port', str(29600 + os.getpid() % 300), os.path.join(ROOT, 'tests', 'dist', script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs at least two GPUs')


@needs2
def test_sharded_render_and_alpha_lattice_equal_single_process():
    out = _torchrun('dist_check.py')
    assert out.count('True; sharded alpha lattice == single-process: True') == 2


@needs2
def test_data_parallel_training_equals_single_process_global_batch():
    """reconstruction() on 2 GPUs x 512 rays == one process with 1024 rays (PSNR to 0.01 dB), incl. the overlapped two-phase
    all-reduce of TrainStep."""
    out = _torchrun('dist_check_train.py')
    assert 'OK' in out


@needs2
def test_image_set_sharded_by_image_equals_union_batch():
    out = _torchrun('dist_check_imageset.py')
    assert 'OK' in out or 'True' in out


@needs2
def test_symmetric_memory_allreduce_kernel_equals_nccl():
    """csrc/allreduce.cu (multimem through NVSwitch, and the peer load/store path) against ncclAllReduce on the 21 MB arena:
    same sums to fp32 rounding, bit-identical on every rank."""
    out = _torchrun('dist_check_allreduce.py')
    assert out.count(': OK') == 2, out[-2000:]
