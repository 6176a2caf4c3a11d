#!/bin/bash
# two GPUs: the multi-GPU part of the suite (incl. peer halos), strong-scaling bench lines with both halo paths, the weak line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_zz_slab_steps.py -q -m gpu > gpurun_out/r2_pytest_multi_peer.log 2>&1
echo "pytest multi rc=$?"; tail -4 gpurun_out/r2_pytest_multi_peer.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for halo in nccl peer; do
  timeout 300 $TR --master-port 2961$((RANDOM % 10)) bench.py --gpus 2 --strong --halo $halo --no-e2e --steps 20 --warmup 5 > gpurun_out/r2_strong_c3_n2_$halo.json 2> gpurun_out/r2_strong_c3_n2_$halo.err
  echo "strong $halo rc=$?"; python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2_strong_c3_n2_$halo.json') if l.startswith('{')][-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'halo', 'parity', 'host_issue_ms_per_step', 'gpu_launches')}, d['roofline']['forward_ms'], d['roofline']['adjoint_ms'])
except Exception as e:
    print('no line', e)
PY
done
timeout 400 $TR --master-port 2962$((RANDOM % 10)) bench.py --gpus 2 --no-e2e --steps 20 --warmup 5 > gpurun_out/r2_weak_c3_n2_auto.json 2> gpurun_out/r2_weak_c3_n2_auto.err
echo "weak auto rc=$?"; grep -o '"ms_per_step": [0-9.]*\|"halo": {[^}]*}\|"sharded_equals_unsharded": [a-z]*' gpurun_out/r2_weak_c3_n2_auto.json | head -5
