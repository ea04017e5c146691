#!/bin/bash
# usage: gpurun --timeout 2400 -- 'bash tools/capture_all.sh <tag>'   (tests + breakdown + bench, then the ncu capture)
T=$1
bash tools/capture_r2l.sh $T
bash tools/capture_prof.sh ${T}p
