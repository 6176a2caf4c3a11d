#!/bin/bash
# peer halos at N GPUs (default 2): correctness (bit-exact against the unsharded kernels) + timing against the NCCL path
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # tag, env, args
  timeout 420 env $2 $TR --master-port 2954$((RANDOM % 10)) scripts/check_peer_halo.py $3 > gpurun_out/r2_peer_n${N}_$1.log 2>&1
  echo "== $1 ($2 $3) rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM\|NCCL version\|^$" gpurun_out/r2_peer_n${N}_$1.log | tail -8
}
run c3_zeros X=1 "c3 zeros 9 --time"
run c3_none X=1 "c3 none 11"
run c4_zeros X=1 "c4 zeros 9 --time"
run c4_none X=1 "c4 none 7"
run c3_thin PSAD_CHECK_SHAPE=128,1024,1024 "c3 zeros 4 --time"
run c3_thin_memset "PSAD_CHECK_SHAPE=128,1024,1024 PSAD_PEER_STREAM_SIGNAL=1" "c3 zeros 4 --time"
run c4_thin PSAD_CHECK_SHAPE=96,768,768 "c4 zeros 4 --time"
