#!/bin/bash
# ncu --set full of one forward+backward at a given shape:  tools/gpu_prof.sh <tag> H W C G [dtype]
tag=$1; shift
dt=${5:-f32}
ncu --set full --clock-control none --import-source on -k regex:'fwd_tiled|bwd_gather|bwd_scatter|merge_far|redo_hot' -c 5 \
    -f -o gpurun_out/${tag} python tools/prof_one.py $1 $2 $3 $4 16 $dt 1 > gpurun_out/${tag}.log 2>&1
tail -3 gpurun_out/${tag}.log
