#!/bin/bash
# A/B timing of library variants built under build/variants/ (run on the GPU box):
#   tools/ab_run.sh "H W C G" ...   -> per variant: ncu durations of the backward kernels
for lib in build/variants/*.so; do
  cp "$lib" iseg_b200/lib/libdcnv3_b200.so
  for s in "$@"; do
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bwd_gather|bwd_scatter|fwd_tiled" --csv --log-file /tmp/ab.csv \
        python tools/prof_one.py $s 16 f32 1 >/dev/null 2>&1
    echo "$(basename $lib) [$s] $(grep -E 'bwd_gather|bwd_scatter|fwd_tiled' /tmp/ab.csv | awk -F'","' '{printf "%s ", $NF}' | tr -d '"')"
  done
done
