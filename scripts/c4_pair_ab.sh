#!/bin/bash
# A/B: lane-ordered pair loads (no shared-memory bank conflicts, +40 selects per step) against plain LDS.128 for the 27-point fp64 kernel
for tune in "lds_pair=0" "lds_pair=1"; do
  echo "== PSAD_TUNE=$tune (burst: 10 launches)"
  PSAD_MARCH_ONLY=1 PSAD_TUNE="$tune" python scripts/kbench.py c4 2>&1 | grep -E "march|rror" | head -3
  echo "== PSAD_TUNE=$tune (sustained: 500 launches)"
  PSAD_TUNE="$tune" python scripts/sustain_trace.py c4 500 2>&1 | cut -c1-150
done
