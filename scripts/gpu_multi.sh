#!/bin/bash
# N GPUs: bench (games sharded, no collective) and a short PPO run (NCCL flat-bucket all-reduce)
N=${1:-2}
mkdir -p gpurun_out
if [ "${BENCH:-1}" = "1" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 500 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err
fi
if [ "${PPO:-1}" = "1" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 scripts/train_ppo.py --games ${GAMES:-65536} --horizon 64 --updates 4 --eval-every 4 --eval-games 2048 > gpurun_out/train_ppo_${N}gpu.log 2>&1; echo "ppo rc=$?"; grep update gpurun_out/train_ppo_${N}gpu.log | cut -c1-420 | tail -4
fi
