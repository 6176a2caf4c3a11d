#!/bin/bash
# What the driver runs at round end, on one GPU, on the tree as it is: the whole GPU suite (no -x), smoke(), the default
# bench line and the reference arm.  Logs / lines land in gpurun_out/r2_final_*.
#   gpurun --timeout 900 -- 'bash scripts/gpu_r2_final_1gpu.sh'
set -u
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_final_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2_final_smoke.log
timeout 400 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2_final_bench_n1.json
if [ "${1:-}" != "noref" ]; then
  timeout 200 python bench.py --impl reference > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err
  echo "reference arm rc=$?"; tail -c 400 gpurun_out/r2_final_bench_ref.json
fi
