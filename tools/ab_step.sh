#!/bin/bash
# A/B of library variants (build/variants/*.so) on the headline step itself: bench.py's CUDA graph of all 30 layers,
# DRAM-cold inputs (each layer owns its tensors) -- unlike tools/ab_run.sh, whose single-layer runs read L2-warm data.
#   tools/ab_step.sh [extra bench.py args]
cp iseg_b200/lib/libdcnv3_b200.so /tmp/keep.so
for lib in build/variants/*.so; do
  cp "$lib" iseg_b200/lib/libdcnv3_b200.so
  echo "$(basename $lib) $(python bench.py --steps 20 --warmup 3 --kernels-only "$@" 2>/dev/null | tail -1)"
done
cp /tmp/keep.so iseg_b200/lib/libdcnv3_b200.so
