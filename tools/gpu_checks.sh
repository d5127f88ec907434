#!/bin/bash
# Correctness checks beyond pytest (run on the GPU box): compute-sanitizer memcheck / racecheck on a few shapes that
# exercise every kernel variant, then the randomized parity fuzzer.   tools/gpu_checks.sh [tag] [fuzz seconds]
tag=${1:-r02}; secs=${2:-150}
out=gpurun_out/${tag}_checks.txt; : > $out
run() { echo "== $*" >> $out; "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|fwd .* us|Error|error" | tail -4 >> $out; }
for s in "49 97 64 4 2 f32" "33 40 80 5 2 f32" "16 16 64 4 3 f32" "49 97 64 4 2 bf16"; do
  run compute-sanitizer --tool memcheck python tools/prof_one.py $s 1
done
OFFSET_SCALE=2 run compute-sanitizer --tool memcheck python tools/prof_one.py 40 40 160 10 2 f32 1
for s in "49 97 64 4 2 f32" "16 16 64 4 3 f32" "40 33 64 4 2 bf16"; do
  run compute-sanitizer --tool racecheck python tools/prof_one.py $s 1
done
echo "== compute-sanitizer --tool memcheck pytest tests/test_gpu_deform_attn.py -k golden" >> $out
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_deform_attn.py -q -k golden 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | tail -2 >> $out
echo "== fuzz" >> $out
timeout $((secs + 120)) python tools/fuzz_gpu.py $secs 11 2>&1 | tail -3 >> $out
cat $out
