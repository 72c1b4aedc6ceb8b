#!/bin/bash
# Refresh of the round-2 record after the last change to the train step (sparse gradient hand-off): full GPU tests, the bench lines
# whose step changed, the step's ncu capture and launch lists, sanitizer.  (The preset / regression ncu captures, the eval line and
# the probes of scratch/r2_final.sh are unaffected and kept.)
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
T=r2f
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gpu_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc $?"
for w in nerf_vm nerf_cp; do timeout 500 python bench.py --workload $w > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; done
for w in image sdf image_set; do timeout 300 python bench.py --workload $w > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fast_fwd_kernel|fast_bwd_saved_agg|mlp2p_fwd|mlp2p_bwd|rgb_fwd_kernel|rgb_bwd_kernel' -s 36 -c 6 -f -o gpurun_out/${T}_prof_step python scratch/prof_step.py > gpurun_out/${T}_ncu_step.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cuda-eager-baseline --eager > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/${T}_launches_graph.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cuda-eager-baseline > gpurun_out/${T}_ncu_launch_graph.log 2>&1
for tool in memcheck racecheck initcheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 1 python scratch/sanitize_case.py > gpurun_out/${T}_sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc $?" | tee -a gpurun_out/${T}_sanitizer_$tool.log
done
du -sh gpurun_out; tail -3 gpurun_out/${T}_gpu_tests.log
