#!/usr/bin/env python3
"""Full-size equivalence check (run on the GPU box): BASELINE configs[2] with the exact threshold
pruning on vs off, and the three kernel paths against each other -- every unit, hit and merged
candidate must be identical (sizes the CPU oracle cannot reach)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from phylign_b200 import _lib
from phylign_b200.matcher import Matcher

n_idx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
args = bench.argparse.Namespace(workload="reads1k", db_scale=1.0, indexes=n_idx, docs=4000, genome_len=1_000_000,
                                reads=100_000, read_len=1000)
w = bench.workload(args)
m = Matcher(0)
specs = [_lib.SynthSpec(**bench.spec_kwargs(b)) for b in w["batches"]]
for b in w["batches"]:
    m.add_synth_index(b["name"], _lib.SynthSpec(**bench.spec_kwargs(b)), b["signature_size"])
m.set_ranks([b["name"] for b in w["batches"]])
raw = m.synth_reads(specs, 3, 0, w["n_reads"], 1000, 51, 655)
m.set_queries_raw(raw, np.arange(w["n_reads"] + 1, dtype=np.uint64) * 1000)


def digest(prune, path):
    m.set_option("prune", prune)
    m.set_option("kernel_path", path)
    m.match_run(0.7, 100, merge_top_n=100)
    res = m.fetch()
    offs, cands = m.merged()
    h = hashlib.sha256()
    # hit offsets depend on allocation order: hash per-unit content in (index, query) order
    u = res.units
    h.update(np.stack([u["query"], u["index"], u["n_pass"], u["n_kept"]]).tobytes())   # not "offset"
    starts = u["offset"].astype(np.int64)
    idx = np.concatenate([np.arange(s, s + n) for s, n in zip(starts, u["n_kept"].astype(np.int64))]) if len(u) else np.zeros(0, np.int64)
    h.update(np.ascontiguousarray(res.hits[idx]).tobytes())
    h.update(offs.tobytes())
    h.update(cands.tobytes())
    return h.hexdigest()[:16], len(u), len(res.hits), len(cands), round(m.phase_ms()[1], 1), m.gathered_bytes()


print(f"# {n_idx} indexes x 4000 docs, 100000 x 1000 bp reads, -t 0.7, top-100; sha256 of units+hits+merged")
ref = None
for prune, path in ((1, 3), (1, 3), (0, 3), (0, 2), (0, 1)):
    d = digest(prune, path)
    ref = ref or d[0]
    print(f"prune={prune} kernel_path={path}: digest {d[0]} units {d[1]} hits {d[2]} merged {d[3]} gather_ms {d[4]} "
          f"gathered_bytes {d[5]}  {'IDENTICAL' if d[0] == ref else 'DIFFERENT'}", flush=True)
    assert d[0] == ref
