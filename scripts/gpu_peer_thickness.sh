#!/bin/bash
# peer halos against the NCCL exchange over slab thickness at N GPUs (C3 fp32 / C4 fp64), timed from an idle GPU in rotating order
# usage: gpu_peer_thickness.sh N "c3 planes..." "c4 planes..."
N=${1:-2}
C3=${2:-"128 256 512 1024"}
C4=${3:-"96 192 384 768"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # tag, env, args
  timeout 420 env $2 $TR --master-port 2954$((RANDOM % 10)) scripts/check_peer_halo.py $3 > gpurun_out/r2_peerthk_n${N}_$1.log 2>&1
  echo "== $1 ($2 $3) rc=$?"; grep "^{\|DIFFERENT\|Error" gpurun_out/r2_peerthk_n${N}_$1.log | tail -3
}
for z in $C3; do run c3_$z PSAD_CHECK_SHAPE=$z,1024,1024 "c3 zeros 3 --time"; done
for z in $C4; do run c4_$z PSAD_CHECK_SHAPE=$z,768,768 "c4 zeros 3 --time"; done
