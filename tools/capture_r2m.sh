#!/bin/bash
# k_march source-level profile (all lines) on the cornell 1024x1024x4 frame, levels 0-2.  usage: capture_r2m.sh <tag>
set -u
T=${1:-r2m}
O=gpurun_out
mkdir -p $O
export RT_B200_LANES=1
python tools/profile_frame.py --size 1024 1024 --spp 4 > $O/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_march' -c 3 -f -o /tmp/${T}_prof \
    python tools/profile_frame.py --size 1024 1024 --spp 4 > $O/${T}_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py /tmp/${T}_prof.ncu-rep > $O/${T}_ncu_march.md 2>&1
for L in 0 1 2; do NCU_ALL_LINES=1 python tools/ncu_source_hotspots.py /tmp/${T}_prof.ncu-rep k_march $L > $O/${T}_march_lines_L$L.txt 2>&1; done
cat $O/${T}_plain.log
