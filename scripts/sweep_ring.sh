#!/usr/bin/env bash
# Tuning sweep of the cp.async ring kernel (run on the GPU box): ring depth x warps per CTA.
set -u
mkdir -p gpurun_out
for cfg in "3 4" "4 4" "2 4" "6 4" "3 8" "4 2" "6 2" "5 3"; do
  set -- $cfg
  PHY_NVCC_DEFS="-DPHY_RING_NB=$1 -DPHY_RING_WARPS=$2" python -m phylign_b200.build --force >/dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  for d in 4000 1000 256; do
    out=$(timeout 300 python bench.py --indexes 8 --docs $d --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1)
    echo "NB=$1 WARPS=$2 docs=$d $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("gather_ms", d["config"]["phase_ms_hash_gather_merge"][1], "frac", round(d["roofline"]["frac"],4))' 2>&1)"
  done
done | tee gpurun_out/sweep_ring.txt
python -m phylign_b200.build --force >/dev/null 2>&1
