#!/bin/bash
# ncu --set full of k_extend / k_march / k_shade, bounce levels 0-2, on the cornell 1024x1024x4 frame (one lane, one batch),
# after the same command has exited 0 without ncu; summaries + per-line hotspots + ncu_kernels.json into gpurun_out/.
# usage: gpurun --timeout 900 -- 'bash tools/capture_prof.sh <tag>'
set -u
T=$1
O=gpurun_out
mkdir -p $O
timeout 200 python tools/kernel_breakdown.py --cfg 3 4b 5 > $O/${T}_breakdown.md 2> $O/${T}_breakdown.err; echo "breakdown rc=$?"; cut -d'|' -f2-9,13- $O/${T}_breakdown.md
export RT_B200_LANES=1
CMD="python tools/profile_frame.py --size 1024 1024 --spp 4"
timeout 120 $CMD > $O/${T}_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_march|k_shade' -c 12 -f -o $O/${T}_prof $CMD > $O/${T}_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py $O/${T}_prof.ncu-rep > $O/${T}_ncu_full.md 2>&1
python tools/ncu_kernels_json.py $T "RT_B200_LANES=1 $CMD" $O/${T}_prof.ncu-rep > $O/${T}_ncu_kernels.json 2> $O/${T}_ncu_kernels.err
REP=$PWD/$O/${T}_prof.ncu-rep   # (absolute: the hotspot tool runs ncu from /tmp)
for K in "k_march<" k_extend k_shade; do NCU_ALL_LINES=1 python tools/ncu_source_hotspots.py $REP "$K" 1 > $O/${T}_${K%<}_lines_L1.txt 2>&1; done
NCU_ALL_LINES=1 python tools/ncu_source_hotspots.py $REP "k_march<" 0 > $O/${T}_k_march_lines_L0.txt 2>&1
ls -la $O/${T}_prof.ncu-rep; cat $O/${T}_plain.log
# gpurun brings back at most 64 MiB: the report itself (~50 MB with sources) stays behind unless asked for
[ "${KEEP_REP:-0}" = "1" ] || rm -f $O/${T}_prof.ncu-rep
