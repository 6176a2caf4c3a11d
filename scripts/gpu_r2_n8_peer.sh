#!/bin/bash
# N GPUs (default 8): peer halos at scale — correctness with two neighbours per rank, the strong-scaling bench line with both
# halo paths, the slab time loop on the strong-scaling slabs with both.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29631 scripts/check_peer_halo.py c3 zeros 9 > gpurun_out/r2_peer_n${N}_check_c3.log 2>&1; echo "check c3 rc=$? identical=$(grep -o IDENTICAL gpurun_out/r2_peer_n${N}_check_c3.log | wc -l) different=$(grep -c DIFFERENT gpurun_out/r2_peer_n${N}_check_c3.log)"
for halo in peer nccl; do
  timeout 300 $TR --master-port 2964$((RANDOM % 10)) bench.py --gpus $N --strong --halo $halo --no-e2e --steps 20 --warmup 5 > gpurun_out/r2_strong_c3_n${N}_$halo.json 2> gpurun_out/r2_strong_c3_n${N}_$halo.err
  echo "strong $halo rc=$?"; grep -o '"ms_per_step": [0-9.]*\|"halo": {[^}]*}\|"sharded_equals_unsharded": [a-z]*\|"forward_ms": [0-9.]*\|"host_issue_ms_per_step": [0-9.]*' gpurun_out/r2_strong_c3_n${N}_$halo.json | head -6 | tr '\n' ' '; echo
done
for mode in "--peer" ""; do
  timeout 300 $TR --master-port 2965$((RANDOM % 10)) scripts/slab_steps_bench.py c3 8 --strong $mode > gpurun_out/r2_slab_steps_strong_c3_n${N}${mode}.json 2> gpurun_out/r2_slab_steps_strong_c3_n${N}${mode}.err
  echo "slab steps strong $mode rc=$?"; grep "^{" gpurun_out/r2_slab_steps_strong_c3_n${N}${mode}.json | cut -c1-400
done
