#!/bin/bash
# k_march3 (per-warp pools): why is it slower?  ncu --set full with source lines, and pool-size variants
set -u
O=gpurun_out
mkdir -p $O
export RT_B200_LANES=1
python tools/profile_frame.py --size 1024 1024 --spp 4 > $O/r2e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_march3' -c 3 -f -o /tmp/r2e_prof \
    python tools/profile_frame.py --size 1024 1024 --spp 4 > $O/r2e_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py /tmp/r2e_prof.ncu-rep > $O/r2e_ncu_march3.md 2>&1
for L in 0 1; do NCU_ALL_LINES=1 python tools/ncu_source_hotspots.py /tmp/r2e_prof.ncu-rep k_march3 $L > $O/r2e_march3_lines_L$L.txt 2>&1; done
unset RT_B200_LANES
for V in "-DRT_M3_R=32" "-DRT_M3_R=96 -DRT_M3_WARPS=2 -DRT_M3_MIN_BLOCKS=3" "-DRT_M3_R=64 -DRT_M3_MIN_BLOCKS=2"; do
  RT_B200_NVCC_EXTRA="$V" python -m rs_pathtracing_b200.build > /dev/null 2>&1
  echo "variant $V"; grep -c "spill" rs_pathtracing_b200/build_rt_march3.log
  timeout 300 python tools/kernel_breakdown.py --cfg 3 2>&1 | tail -1 | cut -c1-120
done
