#!/bin/bash
# one GPU: the time-loop tests (fused pairs through the reference's TimeLoop API), the bench legs that time them, and the
# fused-pair time loop of the fp64 27-point stencil on a one-rank slab (decides whether `_pairs_pay_off` takes fp64)
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_api.py tests/test_gpu_steps.py tests/test_gpu_zz_slab_steps.py -q -m gpu -k "timeloop or steps" > gpurun_out/r2_timeloop_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_timeloop_pytest.log
timeout 120 python scripts/slab_steps_bench.py c4 8 > gpurun_out/r2_slab_steps_c4_n1.json 2> gpurun_out/r2_slab_steps_c4_n1.err
echo "slab steps c4 rc=$?"; cat gpurun_out/r2_slab_steps_c4_n1.json
for w in c3 c2; do
  timeout 200 python bench.py --workload $w --only-headline --no-e2e --no-cpu-baseline > gpurun_out/r2_timeloop_bench_$w.json 2> gpurun_out/r2_timeloop_bench_$w.err
  echo "bench $w rc=$?"; grep -o '"fused_steps": {.*' gpurun_out/r2_timeloop_bench_$w.json | cut -c1-1200
done
