#!/bin/bash
# One-GPU evidence of a round (run on the GPU box): tests, headline bench (both arms), ncu launch list of the step,
# ncu --set full of the three kernels at the largest and the most frequent layer shape (fp32 and bf16), configs 3-5.
#   tools/gpu_evidence.sh [tag]
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --one-dtype --kernels-only > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
bash tools/gpu_prof.sh ${tag}_s1_f32 128 128 64 4 f32
bash tools/gpu_prof.sh ${tag}_s3_f32 32 32 256 16 f32
bash tools/gpu_prof.sh ${tag}_s1_bf16 128 128 64 4 bf16
bash tools/gpu_prof.sh ${tag}_s3_bf16 32 32 256 16 bf16
python tools/config_bench.py cfg3 cfg4 cfg5 siblings > gpurun_out/${tag}_config_bench.jsonl 2>/dev/null; echo "config rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "frac", d["step_hbm"]["frac_of_peak"], "bf16", d.get("bf16", {}).get("ms_per_step"), "e2e_ms", d["e2e"].get("ms_per_step"))
r = json.loads(open("gpurun_out/${tag}_bench_reference.json").read().strip().splitlines()[-1])
print("reference", r.get("value"), r.get("ms_per_step"))
PY
