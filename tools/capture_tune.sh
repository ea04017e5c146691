#!/bin/bash
set -u
for t in 8,8,8 8,16,8 8,24,8 16,16,8 8,16,16 4,12,8 8,12,4 12,20,8; do
  echo "tune $t"; RT_B200_MARCH_TUNE=$t timeout 100 python tools/kernel_breakdown.py --cfg 3 5 2>/dev/null | cut -d'|' -f2-9 | tail -2
done
