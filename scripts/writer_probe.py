#!/usr/bin/env python3
"""Host-only timing of the native match-file writer on a config-3-shaped fake result
(100k queries x 64 indexes, ~78k non-empty units, ~2.2M hit lines).  No GPU needed."""
import ctypes as C, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phylign_b200 import _lib
from phylign_b200.cobs_index import ClassicHeader
from phylign_b200.cobs_text import _cat
from phylign_b200.match_files import MatchFileSet
from phylign_b200.matcher import HIT_DT, UNIT_DT, ResidentIndex

nq, n_idx, n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000, 64, 4000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(1)
n_units = 78_000
cells = np.sort(rng.choice(nq * n_idx, n_units, replace=False))
units = np.zeros(n_units, UNIT_DT)
units["index"], units["query"] = cells // nq, cells % nq
kept = rng.integers(1, 58, n_units)
units["n_kept"], units["n_pass"] = kept, kept + 3
units["offset"] = np.concatenate(([0], np.cumsum(kept)[:-1]))
hits = np.zeros(int(kept.sum()), HIT_DT)
hits["doc"], hits["score"] = rng.integers(0, n_docs, len(hits)), rng.integers(679, 971, len(hits))
nk = np.zeros(nq, np.uint32)
r = _lib.Results(nq, n_idx, n_units, C.cast(units.ctypes.data, C.POINTER(_lib.Unit)), len(hits),
                 C.cast(hits.ctypes.data, C.POINTER(_lib.Hit)), C.cast(nk.ctypes.data, C.POINTER(C.c_uint32)), 0, 0)
names = [f"r{d:06d}_SYN{d:06d}" for d in range(n_docs)]
ixs = {i: ResidentIndex(i, f"synth_species_{i:03d}__01", ClassicHeader(31, 1, n_docs, 10, 1, names)) for i in range(n_idx)}
hcat, hoffs = _cat([f"read_{q:07d}" for q in range(nq)])
td = tempfile.mkdtemp(dir="/dev/shm")
for rep in range(3):
    fs = MatchFileSet({i: os.path.join(td, f"{ixs[i].batch}____q.gz") for i in ixs}, ixs, threads=threads)
    t0 = time.perf_counter()
    fs.write_block(hcat, hoffs, C.pointer(r))
    fs.commit()
    dt = time.perf_counter() - t0
    print(f"{dt*1e3:.1f} ms wall", fs.stats_dict())
