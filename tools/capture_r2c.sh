#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_source_kat.py tests/test_gpu_fullsize.py tests/test_gpu_render.py -m gpu -q -x --timeout 900 > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r2c_pytest.log
python tools/pixel_allowance_probe.py > $O/r2c_pixel_allowance.jsonl 2> $O/r2c_pixel_allowance.err; echo "probe rc=$?"; grep -c pixel $O/r2c_pixel_allowance.jsonl; grep scene $O/r2c_pixel_allowance.jsonl | cut -c1-220
