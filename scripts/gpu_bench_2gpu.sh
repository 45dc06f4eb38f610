#!/bin/bash
# two GPUs under torchrun: the driver's launch line, own arm then reference arm
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/bench_r2af_2gpu.json 2> gpurun_out/bench_r2af_2gpu.err
tail -c 400 gpurun_out/bench_r2af_2gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2af_2gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"])
for k, v in d.get("extra", {}).items():
    print(k, json.dumps(v)[:260])
PY
