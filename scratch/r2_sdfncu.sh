#!/bin/bash
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'fast_fwd|fast_bwd|wide_' -s 30 -c 2 -f -o gpurun_out/r2n_prof_sdf python bench.py --workload sdf --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2n_ncu_sdf.log 2>&1; tail -2 gpurun_out/r2n_ncu_sdf.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'fast_fwd|fast_bwd|wide_' -s 30 -c 2 -f -o gpurun_out/r2n_prof_image python bench.py --workload image --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2n_ncu_image.log 2>&1; tail -2 gpurun_out/r2n_ncu_image.log
