#!/bin/bash
# C4 geometry sweep: 2-cell strips (conflict-free LDS.128) against the shipped 4-cell strips
cd /root/repo
for tune in "" "sx=2,ry=6,ty=42" "sx=2,ry=6,ty=42,shuffle=0" "sx=2,ry=4,ty=28" "sx=2,ry=5,ty=35" "sx=2,ry=8,ty=56" "sx=2,ry=6,ty=42,lookahead=5" "sx=2,ry=6,ty=42,lookahead=3" "sx=2,ry=4,ty=60" "sx=2,ry=3,ty=45" "sx=4,ry=3,ty=21,lookahead=5" "sx=4,ry=2,ty=14" "sx=4,ry=4,ty=28"; do
  echo "== PSAD_TUNE=$tune"
  PSAD_MARCH_ONLY=1 PSAD_TUNE="$tune" timeout 300 python scripts/kbench.py c4 2>&1 | grep -E "march|Error|error" | head -4
done
