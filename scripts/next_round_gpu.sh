#!/bin/bash
# Everything queued for the first GPU call of the next round, in order of value (one box, one GPU; ~12 min):
#   gpurun --timeout 1500 -- 'bash scripts/next_round_gpu.sh'
# and, separately, the two-GPU checks (charged 2x):
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/next_round_gpu.sh multi 2'
# Outputs land in gpurun_out/r2_*.  Nothing here runs under a profiler except the explicit ncu step at the end.
set -u
mkdir -p gpurun_out
if [ "${1:-single}" = "multi" ]; then
  N=${2:-2}
  python -m pytest tests/test_gpu_multi.py tests/test_gpu_zz_slab_steps.py -m gpu -x -q > gpurun_out/r2_pytest_multi.log 2>&1
  echo "pytest multi rc=$?"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      scripts/slab_steps_bench.py c3 8 > gpurun_out/r2_slab_steps_c3_n$N.json 2> gpurun_out/r2_slab_steps_c3_n$N.err
  echo "slab steps rc=$?"; cat gpurun_out/r2_slab_steps_c3_n$N.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
      bench.py --gpus $N --steps 50 --warmup 5 --strong > gpurun_out/r2_strong_c3_n$N.json 2> gpurun_out/r2_strong_c3_n$N.err
  echo "strong rc=$?"; cat gpurun_out/r2_strong_c3_n$N.json
  exit 0
fi
# 1. the whole GPU suite (includes the CPU-replayed-only paths of round 1: slab fused steps, 2-D lifted pairs)
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
#    the files added without a GPU at hand, without -x: every failure at once
python -m pytest tests/test_gpu_zz_golden.py tests/test_gpu_zz_slab_steps.py -m gpu -q > gpurun_out/r2_pytest_gpu_new.log 2>&1
echo "pytest (new files) rc=$?"; tail -15 gpurun_out/r2_pytest_gpu_new.log
# 2. headline bench lines
for w in c3 c4 c2 c5; do
  python bench.py --workload $w --steps 50 --warmup 5 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo "bench $w rc=$?"
done
# 3. fused-pair candidates: two CTAs per SM so that one CTA's per-plane barrier does not idle the SM
#    (all of these emit, compile without local memory beyond a few spill bytes, and sit in the in-tree cubin cache)
python scripts/steps_bench.py \
    c3 "0,0,0,0;2,14,4,0,1,2;2,18,4,0,1,2;2,22,4,0,1,2;4,20,4,0,1,2;4,28,4,0,1,2;2,30,4,0,1;3,33,4,0,1;4,36,4,0,1;4,40,4,0,1;4,44,4,2,1;4,44,4,4,1" \
    c4 "0,0,0,0;2,14,2,0,1,2;2,14,2,0,0,2;2,10,2,0,1,3;2,18,2,0,1,2;2,22,2,3,0;2,22,2,5,0" \
    c2 "0,0,0,0;2,30,4,0,1;2,14,4,0,1,2" > gpurun_out/r2_steps_bench.log 2>&1; echo "steps_bench rc=$?"
# 4. time loop on one slab with ghost planes: ranged fused launches against ranged single steps
python scripts/slab_steps_bench.py c3 8 > gpurun_out/r2_slab_steps_c3_n1.json 2> gpurun_out/r2_slab_steps_c3_n1.err; echo "slab steps rc=$?"
# 4b. the TV adjoint in both adjoint modes (the exact one has 7.5 % fewer instructions in its step loop)
PSAD_MARCH_ONLY=1 python scripts/kbench.py c5 > gpurun_out/r2_kbench_c5_reference.log 2>&1
PSAD_MARCH_ONLY=1 PSAD_ADJOINT_MODE=exact python scripts/kbench.py c5 > gpurun_out/r2_kbench_c5_exact.log 2>&1
echo "kbench c5 rc=$?"; grep backward gpurun_out/r2_kbench_c5_*.log
# 5. ncu of the exchange variant of the fused pair (after its plain run above exited 0)
ncu --set full --clock-control none --import-source on -k regex:march_x2e -c 1 -o gpurun_out/r2_c3_x2e \
    python scripts/steps_bench.py c3 "0,0,0,0" > gpurun_out/r2_ncu_x2e.log 2>&1; echo "ncu rc=$?"
