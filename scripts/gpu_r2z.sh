#!/bin/bash
# K1 variants (build/exp/libbg_PEU.so: Prefetch, Early alloc, Unrolled copy)
cp mlp-ppo-2ply-p3_b200/libbg_b200.so build/exp/libbg_head.so
for v in 000 100 010 001 110 111 000 111; do
  cp build/exp/libbg_$v.so mlp-ppo-2ply-p3_b200/libbg_b200.so
  TAG=$v timeout 200 python scripts/exp_k1_variants.py 2>&1 | tail -1
done
cp build/exp/libbg_head.so mlp-ppo-2ply-p3_b200/libbg_b200.so
