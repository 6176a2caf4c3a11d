#!/bin/bash
# 1 GPU: does the peer-halo instance cost time when there are no neighbours at all? (forward kernel, from idle; PSAD_NVRTC_EXTRA
# selects template variants, e.g. -DPSAD_PEER_PRODUCER_NOINLINE=0 / -DPSAD_PEER_PRODUCER=0)
mkdir -p gpurun_out
i=0
for extra in "" "-DPSAD_PEER_PRODUCER_NOINLINE=0"; do
  i=$((i+1))
  for w in c4 c3; do
  PSAD_NVRTC_EXTRA="$extra" python scripts/peer_codegen_check.py $w --quick > gpurun_out/r2_peer_variant_${w}_$i.json 2> gpurun_out/r2_peer_variant_${w}_$i.err || tail -5 gpurun_out/r2_peer_variant_${w}_$i.err
  cat gpurun_out/r2_peer_variant_${w}_$i.json
  done
done
