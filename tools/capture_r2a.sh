#!/bin/bash
# round 2, first capture: baseline confirmation + per-kernel breakdown of every config + ncu --set full on cfg 4b
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2a_pytest.log
python bench.py > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?"; cut -c1-200 $O/r2a_bench.json
python tools/kernel_breakdown.py > $O/r2a_breakdown.md 2> $O/r2a_breakdown.err; echo "breakdown rc=$?"; cat $O/r2a_breakdown.md
export RT_B200_LANES=1
python tools/profile_frame.py --variant 4b --size 1920 1080 --spp 2 > $O/r2a_plain_4b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_march|k_shade' -c 9 -f -o $O/r2a_prof_4b \
    python tools/profile_frame.py --variant 4b --size 1920 1080 --spp 2 > $O/r2a_ncu_4b.log 2>&1; echo "ncu 4b rc=$?"
python tools/profile_frame.py --scene dupin.json --size 1920 1080 --spp 2 > $O/r2a_plain_dupin.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_march|k_shade' -c 6 -f -o $O/r2a_prof_dupin \
    python tools/profile_frame.py --scene dupin.json --size 1920 1080 --spp 2 > $O/r2a_ncu_dupin.log 2>&1; echo "ncu dupin rc=$?"
ls -la $O | tail -20
