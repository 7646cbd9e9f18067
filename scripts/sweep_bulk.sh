#!/usr/bin/env bash
# Tuning sweep of the bulk gather kernel (run on the GPU box): ring depth x warps per CTA.
set -u
mkdir -p gpurun_out
for cfg in "4 4" "6 4" "3 4" "4 8" "6 2" "8 2" "5 4"; do
  set -- $cfg
  PHY_NVCC_DEFS="-DPHY_BULK_NB=$1 -DPHY_BULK_WARPS=$2" python -m phylign_b200.build --force >/dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  out=$(timeout 300 python bench.py --indexes 8 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1)
  echo "NB=$1 WARPS=$2 $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("gather_ms", d["config"]["phase_ms_hash_gather_merge"][1], "frac", round(d["roofline"]["frac"],4), "value", d["value"])' 2>&1)"
done | tee gpurun_out/sweep_bulk.txt
python -m phylign_b200.build --force >/dev/null 2>&1
