#!/bin/bash
# ncu launch list of one bench step set (our kernels only): per-launch device time, cold-cache and serialised.
# usage: tools/launch_list.sh <out.csv> [bench args...]
out=$1; shift
python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/launch_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:^k_' -s 90 -c 60 --csv --log-file "$out" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/launch_ncu.log 2>&1
