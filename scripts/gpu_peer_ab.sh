#!/bin/bash
# A/B of the peer-halo launch order at N GPUs: rotated z-chunk order (default) against PSAD_PEER_NO_ROTATION=1
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # tag, env, args
  timeout 420 env $2 $TR --master-port 2954$((RANDOM % 10)) scripts/check_peer_halo.py $3 > gpurun_out/r2_peerab_n${N}_$1.log 2>&1
  echo "== $1 ($2 $3) rc=$?"; grep "^{\|DIFFERENT\|Error" gpurun_out/r2_peerab_n${N}_$1.log | tail -3
}
run c3_rot X=1 "c3 zeros 9 --time"
run c3_norot PSAD_PEER_NO_ROTATION=1 "c3 zeros 9 --time"
run c4_rot X=1 "c4 zeros 9 --time"
run c4_norot PSAD_PEER_NO_ROTATION=1 "c4 zeros 9 --time"
run c3_thin_rot PSAD_CHECK_SHAPE=128,1024,1024 "c3 zeros 4 --time"
run c4_thin_rot PSAD_CHECK_SHAPE=96,768,768 "c4 zeros 4 --time"
