#!/bin/bash
# Two GPUs, the tree as it is: the multi-GPU tests (peer halos included; the slab-steps check that failed on the CHECK's side
# in r2_pytest_gpu_2gpu_peer_first.log) and the default weak-scaling bench line exactly as the driver launches it.
#   gpurun --gpus 2 --timeout 600 -- 'bash scripts/gpu_r2_final_2gpu.sh'
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_zz_slab_steps.py::test_fused_steps_sharded_equals_unsharded" -q -m gpu > gpurun_out/r2_final_pytest_2gpu.log 2>&1
echo "pytest multi rc=$?"; tail -4 gpurun_out/r2_final_pytest_2gpu.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n2.json 2> gpurun_out/r2_final_bench_n2.err
echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/r2_final_bench_n2.json
