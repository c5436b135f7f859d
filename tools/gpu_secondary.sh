#!/bin/bash
# secondary workloads of BASELINE.json (configs[2], configs[3], float32 input, sparse air) with the built library
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
for args in "--sps 8" "--density 5e-5" "--f32" "--density 0.05" "--sc16"; do
  timeout 14 python bench.py --steps 40 --warmup 3 --no-cpu --no-e2e $args >> gpurun_out/secondary.jsonl 2>> gpurun_out/secondary.log
  echo "[+$(( $(date +%s) - T0 ))s] $args rc=$?"
done
