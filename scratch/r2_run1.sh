#!/bin/bash
# GPU run 1 of round 2: parity tests, bench (with the reference's CUDA-eager and CPU legs), reference arm, L2-red probe,
# ncu captures of the field / linear_mat kernels, compute-sanitizer over the small cases.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc $?"
timeout 120 python scratch/probe_red.py > gpurun_out/r2_probe_red.json 2> gpurun_out/r2_probe_red.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fast_fwd_kernel|fast_bwd_saved_agg|mlp2_fwd|mlp2_bwd' -s 24 -c 4 -f -o gpurun_out/r2_prof_field python scratch/prof_step.py > gpurun_out/r2_ncu_field.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cuda-eager-baseline --eager > gpurun_out/r2_ncu_launch.log 2>&1
for tool in memcheck racecheck initcheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 1 python scratch/sanitize_case.py > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc $?"
done
tail -5 gpurun_out/r2_gpu_tests.log
cat gpurun_out/r2_bench.json | head -c 3000
