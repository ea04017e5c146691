#!/bin/bash
# every configuration through bench.py, reference arm (CPU) and our arm, one JSON line each
set -u
O=gpurun_out
mkdir -p $O
for c in 1 2 4a 4b 5; do
  timeout 300 python bench.py --impl reference --config $c --steps 2 --warmup 1 > $O/table_ref_$c.json 2> $O/table_ref_$c.err; echo "ref $c rc=$?"; cut -c1-150 $O/table_ref_$c.json
  timeout 300 python bench.py --config $c --no-cpu-baseline > $O/table_ours_$c.json 2> $O/table_ours_$c.err; echo "ours $c rc=$?"; cut -c1-150 $O/table_ours_$c.json
done
