#!/bin/bash
# round 2, capture b: k_march source-level profile (all lines) on the cornell 1024x1024x4 frame, levels 0-2
set -u
O=gpurun_out
mkdir -p $O
python tools/kernel_breakdown.py --cfg 3 5 > $O/r2b_breakdown.md 2> $O/r2b_breakdown.err; echo "breakdown rc=$?"; cat $O/r2b_breakdown.md
export RT_B200_LANES=1
python tools/profile_frame.py --size 1024 1024 --spp 4 > $O/r2b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_march' -c 3 -f -o /tmp/r2b_prof \
    python tools/profile_frame.py --size 1024 1024 --spp 4 > $O/r2b_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py /tmp/r2b_prof.ncu-rep > $O/r2b_ncu_march.md 2>&1
ncu -i /tmp/r2b_prof.ncu-rep --page raw --csv --metrics sm__cycles_active.avg,sm__cycles_elapsed.avg,smsp__cycles_active.avg > $O/r2b_cycles.csv 2>&1
for L in 0 1 2; do NCU_ALL_LINES=1 python tools/ncu_source_hotspots.py /tmp/r2b_prof.ncu-rep k_march $L > $O/r2b_march_lines_L$L.txt 2>&1; done
python tools/profile_frame.py --scene dupin.json --size 1920 1080 --spp 2 > $O/r2b_plain_dupin.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_march' -c 2 -f -o /tmp/r2b_prof_dupin \
    python tools/profile_frame.py --scene dupin.json --size 1920 1080 --spp 2 > $O/r2b_ncu_dupin.log 2>&1; echo "ncu dupin rc=$?"
python tools/summarize_ncu.py /tmp/r2b_prof_dupin.ncu-rep > $O/r2b_ncu_march_dupin.md 2>&1
for L in 0 1; do NCU_ALL_LINES=1 python tools/ncu_source_hotspots.py /tmp/r2b_prof_dupin.ncu-rep k_march $L > $O/r2b_march_dupin_lines_L$L.txt 2>&1; done
ls -la $O | grep r2b
