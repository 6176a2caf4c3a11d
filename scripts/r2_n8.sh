#!/bin/bash
# Round-2 multi-GPU validation: weak-scaling bench line (with sharded == unsharded parity and the chunk-streamed e2e on every
# rank), the strong-scaling line, and the slab time loop with fused pairs.  Usage: bash scripts/r2_n8.sh N
N=${1:-8}
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/r2_n${N}_host.txt; nproc >> gpurun_out/r2_n${N}_host.txt; nvidia-smi topo -m >> gpurun_out/r2_n${N}_host.txt 2>&1
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $R --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "weak rc=$?"
timeout 600 $R --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --strong > gpurun_out/r2_strong_c3_n$N.json 2> gpurun_out/r2_strong_c3_n$N.err; echo "strong rc=$?"
timeout 600 $R --master-port 29523 scripts/slab_steps_bench.py c3 8 > gpurun_out/r2_slab_steps_c3_n$N.json 2> gpurun_out/r2_slab_steps_c3_n$N.err; echo "slab steps rc=$?"
tail -c 400 gpurun_out/r2_bench_n$N.err
