#!/usr/bin/env python3
"""BASELINE configs[4]: long-read shape with the match-ratio threshold swept 0.4 .. 0.9 (run on
the GPU box).  Shows what the exact threshold pruning buys at each -t: rows really gathered and
time per pass, for 10 kbp reads (14 counter planes) and 1 kbp reads (10 planes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from phylign_b200 import _lib
from phylign_b200.matcher import Matcher

n_idx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
args = bench.argparse.Namespace(workload="reads1k", db_scale=1.0, indexes=n_idx, docs=4000, genome_len=1_000_000,
                                reads=10_000, read_len=10_000)
w = bench.workload(args)
m = Matcher(0)
specs = [_lib.SynthSpec(**bench.spec_kwargs(b)) for b in w["batches"]]
for b in w["batches"]:
    m.add_synth_index(b["name"], _lib.SynthSpec(**bench.spec_kwargs(b)), b["signature_size"])
m.set_ranks([b["name"] for b in w["batches"]])
print(f"# {n_idx} indexes x 4000 docs (1.43 GB each); columns: reads x len, -t, gather ms/pass, "
      "row bytes gathered / all-pairs bytes, bases/s, units with hits, hits")
for n_reads, rlen, err in ((10_000, 10_000, 5243), (100_000, 1_000, 655)):
    raw = m.synth_reads(specs, 5, 0, n_reads, rlen, 51, err)      # 8% errors for the nanopore-like reads
    offs = np.arange(n_reads + 1, dtype=np.uint64) * rlen
    m.set_queries_raw(raw, offs)
    all_pairs = n_reads * (rlen - 30) * n_idx * 500
    for thr in (0.4, 0.5, 0.6, 0.7, 0.8, 0.9):
        ms = []
        for it in range(4):
            m.match_run(thr, 100, merge_top_n=100)
            if it:
                ms.append(sum(m.phase_ms()[:3]))
        res = m.fetch()
        g = m.gathered_bytes()
        t = float(np.mean(ms))
        print(f"{n_reads} x {rlen}  -t {thr}  {m.phase_ms()[1]:8.1f} ms  gathered {g / all_pairs:5.3f}  "
              f"{n_reads * rlen / (t * 1e-3):.3e} bases/s  units {len(res.units)}  hits {len(res.hits)}", flush=True)
