#!/bin/bash
# Multi-GPU evidence (run with gpurun --gpus N): bit-exact shard-vs-single test, then the headline bench at N ranks.
n=${1:-2}; tag=${2:-r02}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo_${n}gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs > gpurun_out/${tag}_multi_${n}gpu.log 2>&1
tail -4 gpurun_out/${tag}_multi_${n}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${n}gpu.json").read().strip().splitlines()[-1])
    print("n_gpus", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"])
    print("e2e", json.dumps(d["e2e"]))
    print("gather", json.dumps(d.get("gather")))
except Exception as e:
    print("no bench line:", e); print(open("gpurun_out/${tag}_bench_${n}gpu.err").read()[-2000:])
PY
