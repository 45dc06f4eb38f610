#!/bin/bash
# N GPUs under torchrun, as the driver launches it: bash scripts/gpu_bench_ngpu.sh N   (gpurun --gpus N)
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -c 400 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${N}gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"])
for k, v in d.get("extra", {}).items():
    print(k, json.dumps(v)[:200])
PY
