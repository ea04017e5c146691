#!/bin/bash
# 2 GPUs: the whole GPU suite (multi-device handle, f64 accumulators, raygen fusion), breakdown, bench at N = 1 and 2
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $O/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r2i_pytest.log
timeout 600 python tools/kernel_breakdown.py --cfg 1 3 4a 4b 5 > $O/r2i_breakdown.md 2> $O/r2i_breakdown.err; echo "breakdown rc=$?"; cut -c1-150 $O/r2i_breakdown.md
timeout 600 python bench.py > $O/r2i_bench_n1.json 2> $O/r2i_bench_n1.err; echo "bench rc=$?"; cut -c1-260 $O/r2i_bench_n1.json
timeout 600 python bench.py --gpus 2 > $O/r2i_bench_n2.json 2> $O/r2i_bench_n2.err; echo "bench2 rc=$?"; cut -c1-260 $O/r2i_bench_n2.json
timeout 300 python tools/multi_handle_probe.py > $O/r2i_multi_probe.txt 2>&1; echo "probe rc=$?"; cat $O/r2i_multi_probe.txt
