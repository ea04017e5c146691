#!/bin/bash
# k_shade: register cap restored; binned (material-keyed) variant against arrival order on cfg 3 / 4b / 5; ncu on 4b
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -q -x --timeout 600 > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2j_pytest.log
echo "arrival order"; timeout 600 python tools/kernel_breakdown.py --cfg 3 4b 5 2>&1 | cut -c1-100
echo "binned"; RT_B200_SHADE_BINNED=1 timeout 600 python tools/kernel_breakdown.py --cfg 3 4b 5 2>&1 | cut -c1-100
export RT_B200_LANES=1
for B in 0 1; do
  export RT_B200_SHADE_BINNED=$B
  python tools/profile_frame.py --variant 4b --size 1920 1080 --spp 2 > $O/r2j_plain_4b_$B.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'k_shade' -c 3 -f -o /tmp/r2j_prof_$B \
      python tools/profile_frame.py --variant 4b --size 1920 1080 --spp 2 > $O/r2j_ncu_$B.log 2>&1; echo "ncu $B rc=$?"
  python tools/summarize_ncu.py /tmp/r2j_prof_$B.ncu-rep > $O/r2j_ncu_shade_4b_binned$B.md 2>&1
  grep -n "duration\|active threads\|warp instructions\|issue-slot\|regs" $O/r2j_ncu_shade_4b_binned$B.md
done
