#!/bin/bash
# final single-GPU capture of a round: tests, reference arm, bench, every configuration, launch list, ncu captures
# usage: gpurun --timeout 3000 -- 'bash tools/capture_final.sh <tag>'
set -u
T=$1
O=gpurun_out
mkdir -p $O
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 1000 python -m pytest tests -m gpu -q --timeout 240 > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
fi
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${T}_smoke.log
ONLY_FULL=0 timeout 1500 bash tools/capture.sh $T
python tools/summarize_launches.py $O/launches_$T.csv > $O/${T}_launches_summary.txt 2>&1
rm -f $O/prof_$T.ncu-rep      # (capture_prof.sh makes the report the summaries come from; 64 MiB come back at most)
timeout 900 bash tools/capture_prof.sh $T
