#!/bin/bash
# k_march3 first run: correctness (whole GPU suite, bounded) then timing against k_march
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r2d_pytest.log
timeout 300 python tools/kernel_breakdown.py --cfg 1 3 4b 5 > $O/r2d_breakdown_v3.md 2> $O/r2d_breakdown_v3.err; echo "v3 rc=$?"; cat $O/r2d_breakdown_v3.md
RT_B200_MARCH=1 timeout 300 python tools/kernel_breakdown.py --cfg 3 5 > $O/r2d_breakdown_v1.md 2> $O/r2d_breakdown_v1.err; echo "v1 rc=$?"; cat $O/r2d_breakdown_v1.md
timeout 300 python bench.py --no-cpu-baseline > $O/r2d_bench.json 2> $O/r2d_bench.err; echo "bench rc=$?"; cut -c1-300 $O/r2d_bench.json
