#!/usr/bin/env bash
# Throughput by row-size class (run on the GPU box): same reads, 8 indexes, varying docs/index.
mkdir -p gpurun_out
for d in 100 128 256 400 512 664 1000 1407 2000 3000 4000; do
  out=$(timeout 300 python bench.py --indexes 8 --docs $d --steps 3 --warmup 2 --no-cpu-baseline "$@" 2>&1 | tail -1)
  echo "docs=$d $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("row_B", (d["config"]["docs_per_index"]+7)//8, "gather_ms", d["config"]["phase_ms_hash_gather_merge"][1], "alg_GBps", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "units", d["config"]["n_units"], "hits", d["config"]["n_hits"])' 2>&1)"
done | tee gpurun_out/sweep_docs.txt
