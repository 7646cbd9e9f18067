#!/usr/bin/env python3
"""Per-iteration timing of the host<->device legs of one step (run on the GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from phylign_b200 import _lib
from phylign_b200.matcher import Matcher, PinnedBuffer

n_idx = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
rlen = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
docs = int(sys.argv[4]) if len(sys.argv) > 4 else 4000
args = bench.argparse.Namespace(workload="reads1k", db_scale=1.0, indexes=n_idx, docs=docs, genome_len=1_000_000, reads=n_reads, read_len=rlen)
w = bench.workload(args)
m = Matcher(0)
specs = [_lib.SynthSpec(**bench.spec_kwargs(i, w)) for i in range(n_idx)]
for i in range(n_idx):
    m.add_synth_index(bench.batch_name(i), specs[i], w["signature_size"])
m.set_ranks([bench.batch_name(i) for i in range(n_idx)])
raw = m.synth_reads(specs, 3, 0, w["n_reads"], rlen, 51, 655)
offs = np.arange(w["n_reads"] + 1, dtype=np.uint64) * rlen
pr, po = PinnedBuffer(len(raw)), PinnedBuffer(offs.nbytes)
pr.array[:] = np.frombuffer(raw, dtype=np.uint8); po.array[:] = offs.view(np.uint8)
print(f"# {n_idx} indexes x {docs} docs ({(docs + 7) // 8}-B rows), {n_reads} reads x {rlen} bp; algorithmic bytes per step {w['alg_bytes']}", flush=True)
for it in range(8):
    t = [time.perf_counter()]
    m.set_queries_raw(pr, po); t.append(time.perf_counter())
    m.match_run(0.7, 100, merge_top_n=100); t.append(time.perf_counter())
    res = m.fetch(); t.append(time.perf_counter())
    mo, mc = m.merged(); t.append(time.perf_counter())
    print(it, [round((b - a) * 1e3, 2) for a, b in zip(t, t[1:])], 'device phases', [round(x, 2) for x in m.phase_ms()], len(res.units), len(res.hits), len(mc), 'gathered', m.gathered_bytes(), flush=True)
