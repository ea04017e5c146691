#!/bin/bash
# k_march3 with one pool per SM + per-phase queues: correctness, timing, profile
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python tools/kernel_breakdown.py --cfg 3 5 > $O/r2g_breakdown.md 2> $O/r2g_breakdown.err; echo "breakdown rc=$?"; cut -c1-140 $O/r2g_breakdown.md; tail -3 $O/r2g_breakdown.err
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r2g_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $O/r2g_bench.json 2> $O/r2g_bench.err; echo "bench rc=$?"; cut -c1-200 $O/r2g_bench.json
export RT_B200_LANES=1
python tools/profile_frame.py --size 1024 1024 --spp 4 > $O/r2g_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_march3' -c 2 -f -o /tmp/r2g_prof \
    python tools/profile_frame.py --size 1024 1024 --spp 4 > $O/r2g_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py /tmp/r2g_prof.ncu-rep > $O/r2g_ncu_march3.md 2>&1
for L in 0 1; do NCU_ALL_LINES=1 python tools/ncu_source_hotspots.py /tmp/r2g_prof.ncu-rep k_march3 $L > $O/r2g_march3_lines_L$L.txt 2>&1; done
