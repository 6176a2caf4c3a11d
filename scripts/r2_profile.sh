#!/bin/bash
# Round-2 profiler evidence (one GPU).  Every ncu / sanitizer step runs only after the same command exited 0 without it.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --only-headline --no-cpu-baseline --e2e-steps 1"
$B > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3.csv $B > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:heat3d_forward_gpu_march -c 1 -o gpurun_out/r2_c3_fwd $B > gpurun_out/r2_ncu_full_c3.log 2>&1; echo "ncu full c3 rc=$?"
for w in c4 c2 c5; do
  K="python scripts/kbench.py $w"
  PSAD_MARCH_ONLY=1 $K > gpurun_out/r2_kbench_$w.log 2>&1; echo "kbench $w rc=$?"
  PSAD_MARCH_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:forward_gpu_march -c 1 -o gpurun_out/r2_${w}_fwd $K > gpurun_out/r2_ncu_full_$w.log 2>&1; echo "ncu full $w fwd rc=$?"
done
PSAD_MARCH_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:tvgrad_backward_gpu_march -c 1 -o gpurun_out/r2_c5_bwd python scripts/kbench.py c5 > gpurun_out/r2_ncu_full_c5b.log 2>&1; echo "ncu full c5 bwd rc=$?"
# compute-sanitizer is closed on this pool (rc 86, "runs under it have left GPUs needing a reset"): the small-launch script
# below still runs plain; the race evidence for the row-exchange kernel is the CPU replay with concurrent warps
# (tests/test_march_replay.py::test_replay_detects_a_missing_consumer_barrier)
python scripts/sanitize_small.py > gpurun_out/r2_sanitize_plain.log 2>&1; echo "sanitize plain rc=$?"
