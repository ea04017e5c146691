#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for rep in 1 2; do timeout 45 python tools/hang_probe.py 16 2>&1 | tail -3; echo "probe rep $rep rc=${PIPESTATUS[0]}"; done
timeout 45 python tools/hang_probe.py 2 2>&1 | tail -3; echo "probe spp2 rc=${PIPESTATUS[0]}"
timeout 120 python tools/kernel_breakdown.py --cfg 3 4b 5 > $O/flat_breakdown.md 2> $O/flat_breakdown.err; echo "breakdown rc=$?"; cut -d'|' -f2-9,13- $O/flat_breakdown.md
