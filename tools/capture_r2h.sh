#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for V in "-DRT_M3_THREADS=256" "-DRT_M3_THREADS=192" "-DRT_M3_THREADS=128" "-DRT_M3_THREADS=384"; do
  RT_B200_NVCC_EXTRA="$V" python -m rs_pathtracing_b200.build > /dev/null 2>&1
  echo "variant $V"
  timeout 300 python tools/kernel_breakdown.py --cfg 3 5 2>&1 | tail -2 | cut -c1-110
done
RT_B200_NVCC_EXTRA="-DRT_M3_THREADS=192" python -m rs_pathtracing_b200.build > /dev/null 2>&1
export RT_B200_LANES=1
python tools/profile_frame.py --size 1024 1024 --spp 4 > $O/r2h_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_march3' -c 2 -f -o /tmp/r2h_prof \
    python tools/profile_frame.py --size 1024 1024 --spp 4 > $O/r2h_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py /tmp/r2h_prof.ncu-rep > $O/r2h_ncu_march3.md 2>&1
for L in 1; do NCU_ALL_LINES=1 python tools/ncu_source_hotspots.py /tmp/r2h_prof.ncu-rep k_march3 $L > $O/r2h_march3_lines_L$L.txt 2>&1; done
grep -n "duration\|active threads\|warp instructions\|issue-slot\|IPC" $O/r2h_ncu_march3.md
