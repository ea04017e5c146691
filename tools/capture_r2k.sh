#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r2k_pytest.log
timeout 600 python tools/big_scene_probe.py > $O/r2k_big_scenes.md 2> $O/r2k_big_scenes.err; echo "big rc=$?"; cat $O/r2k_big_scenes.md; tail -3 $O/r2k_big_scenes.err
timeout 600 python tools/kernel_breakdown.py --cfg 1 3 4a 4b 5 > $O/r2k_breakdown.md 2> $O/r2k_breakdown.err; echo "breakdown rc=$?"; cut -c1-110 $O/r2k_breakdown.md
timeout 600 python bench.py --no-cpu-baseline > $O/r2k_bench.json 2> $O/r2k_bench.err; echo "bench rc=$?"; cut -c1-200 $O/r2k_bench.json
