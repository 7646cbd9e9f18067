#!/usr/bin/env python3
"""First-call vs steady-state cost of one match pass (run on the GPU box): wall and device phases of
the first three phy_match_run + fetch + merged calls of a fresh context."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from phylign_b200 import _lib
from phylign_b200.matcher import Matcher

n_idx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rlen = int(sys.argv[2]) if len(sys.argv) > 2 else 150
pinned = int(sys.argv[3]) if len(sys.argv) > 3 else 1
args = bench.argparse.Namespace(workload="reads1k", db_scale=1.0, indexes=n_idx, docs=4000, genome_len=1_000_000, reads=100_000, read_len=rlen)
w = bench.workload(args)
m = Matcher(0)
specs = [_lib.SynthSpec(**bench.spec_kwargs(i, w)) for i in range(n_idx)]
for i in range(n_idx):
    m.add_synth_index(bench.batch_name(i), specs[i], w["signature_size"])
m.set_ranks([bench.batch_name(i) for i in range(n_idx)])
raw = m.synth_reads(specs, 3, 0, w["n_reads"], rlen, 51, 655)
offs = np.arange(w["n_reads"] + 1, dtype=np.uint64) * rlen
m.set_option("pinned_results", pinned)
if len(sys.argv) > 4 and sys.argv[4] == "prewarm":      # a tiny match first: module load + clocks
    t0 = time.perf_counter()
    tiny = m.add_synth_index("tiny__01", _lib.SynthSpec(seed=1, n_docs=4000, genome_len=20000, clade_size=32, clade_sub_q16=328, doc_sub_q16=328), 60000)
    m.set_active_only([tiny])
    m.set_queries_raw(raw[:rlen * 2000], offs[:2001])
    m.match_run(0.7, 100, merge_top_n=100)
    m.evict(tiny)
    m.set_active_only(list(m.indexes))
    print("prewarm", round((time.perf_counter() - t0) * 1e3, 1), "ms", [round(x, 2) for x in m.phase_ms()], flush=True)
for it in range(4):
    t = [time.perf_counter()]
    m.set_queries_raw(raw, offs); t.append(time.perf_counter())
    m.match_run(0.7, 100, merge_top_n=100); t.append(time.perf_counter())
    res = m.fetch(); t.append(time.perf_counter())
    mo, mc = m.merged(); t.append(time.perf_counter())
    print(it, "wall ms set/run/fetch/merged", [round((b - a) * 1e3, 1) for a, b in zip(t, t[1:])], "device phases", [round(x, 2) for x in m.phase_ms()], len(res.units), len(res.hits), len(mc), flush=True)
