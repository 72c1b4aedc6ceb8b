#!/bin/bash
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
for v in 0 1 0 1; do
  FFB_SPARSE_FEAT_GRAD=$v timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k scheduled > gpurun_out/r2p_sched_$v.log 2>&1
  echo "sparse=$v: $(tail -1 gpurun_out/r2p_sched_$v.log)  $(grep -o 'abs((.*' gpurun_out/r2p_sched_$v.log | head -2 | tr '\n' ' ')"
done
