#!/bin/bash
# Round-2 record run: full GPU tests, the bench lines (headline + presets + regression drivers + eval), probes, ncu captures,
# launch list, sanitizer.  Everything lands in gpurun_out/r2f_*; scratch/collect_profiles.py turns it into profiles/r02_*.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
T=r2f
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_gpu_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
for w in nerf_vm nerf_cp; do timeout 500 python bench.py --workload $w > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; done
for w in image sdf image_set nerf_eval; do timeout 300 python bench.py --workload $w > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; done
timeout 120 python scratch/probe_red.py > gpurun_out/${T}_probe_red.json 2>/dev/null
timeout 120 python scratch/probe_mma.py > gpurun_out/${T}_probe_mma.txt 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fast_fwd_kernel|fast_bwd_saved_agg|mlp2p_fwd|mlp2p_bwd|rgb_fwd_kernel|rgb_bwd_kernel' -s 36 -c 6 -f -o gpurun_out/${T}_prof_step python scratch/prof_step.py > gpurun_out/${T}_ncu_step.log 2>&1
# (gpurun brings back at most 64 MiB: the preset / regression captures are summarised on the box and only the step capture travels)
summarise() { python scratch/ncu_summary.py gpurun_out/${T}_prof_$1.ncu-rep gpurun_out/${T}_ncusum_$1 "$2" "$3" > /dev/null 2>&1 && rm -f gpurun_out/${T}_prof_$1.ncu-rep; }
for p in nerf_vm nerf_cp; do
  timeout 400 ncu --set full --clock-control none -k regex:'lines_fwd|lines_bwd|vm_fwd|vm_bwd' -s 4 -c 2 -f -o gpurun_out/${T}_prof_$p python scratch/run_presets.py $p 4 > gpurun_out/${T}_ncu_$p.log 2>&1
  summarise $p "Round 2: field kernels of the ${p#nerf_} preset at the Tanks&Temples bench shape, ncu --set full, one launch each" "ncu --set full --clock-control none -k regex:'lines_fwd|lines_bwd|vm_fwd|vm_bwd' -s 4 -c 2 python scratch/run_presets.py $p 4"
done
for w in image sdf image_set; do
  timeout 400 ncu --set full --clock-control none -k regex:'fast_fwd|fast_bwd|wide_' -s 30 -c 2 -f -o gpurun_out/${T}_prof_$w python bench.py --workload $w --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/${T}_ncu_$w.log 2>&1
  summarise $w "Round 2: field kernels of the $w.yaml regression step, ncu --set full, one launch each" "ncu --set full --clock-control none -k regex:'fast_fwd|fast_bwd|wide_' -s 30 -c 2 python bench.py --workload $w --no-cpu-baseline --steps 5 --warmup 3"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cuda-eager-baseline --eager > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/${T}_launches_graph.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cuda-eager-baseline > gpurun_out/${T}_ncu_launch_graph.log 2>&1
for tool in memcheck racecheck initcheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 1 python scratch/sanitize_case.py > gpurun_out/${T}_sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc $?" | tee -a gpurun_out/${T}_sanitizer_$tool.log
done
du -sh gpurun_out; tail -3 gpurun_out/${T}_gpu_tests.log
