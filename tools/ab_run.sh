#!/bin/bash
# A/B timing of library variants built under build/variants/ (run on the GPU box):
#   tools/ab_run.sh "H W C G [dtype]" ...   -> per variant: ncu durations (us) of the library's kernels, last launch of each
cp iseg_b200/lib/libdcnv3_b200.so /tmp/keep.so
for lib in build/variants/*.so; do
  cp "$lib" iseg_b200/lib/libdcnv3_b200.so
  for s in "$@"; do
    a=($s)
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bwd_gather|bwd_scatter|fwd_tiled|merge_far|redo_hot" --csv --log-file /tmp/ab.csv \
        python tools/prof_one.py ${a[0]} ${a[1]} ${a[2]} ${a[3]} 16 ${a[4]:-f32} 2 >/dev/null 2>&1
    python - "$lib" "$s" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open("/tmp/ab.csv")) if len(r) > 5 and r[0].isdigit()]
last = {}
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("dcnv3::", "").split("<")[0]
    v = float(r[-1].replace(",", "")); u = r[-2]
    last[name] = v / 1e3 if u in ("ns", "nsecond") else v
print(f"{sys.argv[1].split('/')[-1]:28s} [{sys.argv[2]}] " + "  ".join(f"{k.replace('_kernel','')}={v:.1f}" for k, v in last.items()))
PY
  done
done
cp /tmp/keep.so iseg_b200/lib/libdcnv3_b200.so
