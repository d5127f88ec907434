#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, then the headline bench; outputs under gpurun_out/<tag>_*
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("ms_per_step", d["ms_per_step"], "frac", d["step_hbm"]["frac_of_peak"], "bf16", d.get("bf16", {}).get("ms_per_step"), "e2e_ms", d["e2e"].get("ms_per_step"))
    for k in d["kernels"]:
        print(f'{k["kernel"]:34s} {k["avg_us"]:8.1f} us {k["gbs"]:8.1f} GB/s share {k["share"]:.3f}')
except Exception as e:
    print("no bench line:", e)
    print(open("gpurun_out/${tag}_bench.err").read()[-3000:])
PY
