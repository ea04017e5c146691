#!/bin/bash
# One full single-GPU capture into gpurun_out/ (run on the GPU box through gpurun; tools/make_profiles.py <id>
# then copies it into profiles/ and regenerates profiles/README.md):
#   gpurun --timeout 1500 -- 'bash tools/capture.sh r1z'
# Multi-GPU lines are separate calls:  gpurun --gpus N -- 'bash tools/capture.sh r1z N'
# Every ncu pass runs only after the same command has exited 0 without ncu; numbers printed under ncu are
# never quoted as bench values.
set -u
ID=$1
N=${2:-1}
O=gpurun_out
mkdir -p $O
if [ "$N" != "1" ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N > $O/bench_${ID}_n$N.json 2> $O/bench_${ID}_n$N.err
    echo "rc=$? N=$N"; cat $O/bench_${ID}_n$N.json | cut -c1-200
    exit 0
fi
if [ "${ONLY_FULL:-0}" != "1" ]; then
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$ID.json 2> $O/bench_ref_$ID.err; echo "ref rc=$?"
python bench.py > $O/bench_$ID.json 2> $O/bench_$ID.err; echo "bench rc=$?"; cut -c1-160 $O/bench_$ID.json
python tools/run_configs.py > $O/configs_$ID.md 2> $O/configs_$ID.err; echo "configs rc=$?"
python bench.py --steps 1 --warmup 1 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$ID.csv \
    python bench.py --steps 1 --warmup 1 > $O/ncu_bench_$ID.log 2>&1; echo "launch list rc=$?"
fi
# one lane, so that the 1024x1024x4 frame is ONE batch of 4 Mi paths (bounce levels 0-2 = the first 9 launches)
export RT_B200_LANES=1
python tools/profile_frame.py --size 1024 1024 --spp 4 > $O/plain_$ID.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_march|k_shade' -c 9 -f -o $O/prof_$ID \
    python tools/profile_frame.py --size 1024 1024 --spp 4 > $O/ncu_$ID.log 2>&1; echo "ncu full rc=$?"
