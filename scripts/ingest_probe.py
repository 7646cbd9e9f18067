#!/usr/bin/env python3
"""Index ingest throughput (run on the GPU box): xz files -> host decode -> pinned -> HBM.

Builds N synthetic indexes on the device, writes them as .cobs_classic.xz (xz -1 -T0), then
times Matcher.load_index (sequential) against Matcher.load_indexes(workers=W)."""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phylign_b200 import _lib
from phylign_b200.cobs_index import ClassicHeader
from phylign_b200.matcher import Matcher

n_idx, docs, glen = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
td = tempfile.mkdtemp(dir="/tmp")
m = Matcher(0)
paths, total = [], 0
for i in range(n_idx):
    spec = _lib.SynthSpec(seed=70 + i, n_docs=docs, genome_len=glen, clade_size=32, clade_sub_q16=328, doc_sub_q16=328)
    sig = int((glen - 30) * 2.8037) + 1
    idx = m.add_synth_index(f"ing__{i:02d}", spec, sig)
    body = m.download_index(idx)
    hdr = ClassicHeader(31, 1, docs, sig, 1, m.indexes[idx].doc_names)
    p = os.path.join(td, f"ing__{i:02d}.cobs_classic")
    with open(p, "wb") as f:
        f.write(hdr.to_bytes()); f.write(body)
    total += len(body)
    subprocess.check_call(["xz", "-1", "-T0", "-k", "-f", p])
    paths.append(p + ".xz")
    m.evict(idx)
print(f"{n_idx} indexes, {total/1e9:.2f} GB decompressed, {sum(os.path.getsize(p) for p in paths)/1e9:.2f} GB xz", flush=True)
t0 = time.perf_counter(); ids = [m.load_index(p) for p in paths[:2]]; m.sync(); t1 = time.perf_counter(); t1 = t0 + (t1 - t0) * len(paths) / 2; ids += m.load_indexes(paths[2:], workers=16)
print(f"sequential load_index: {t1-t0:.2f} s = {total/(t1-t0)/1e9:.2f} GB/s", flush=True)
chk = m.download_index(ids[0])[:64]
for i in ids: m.evict(i)
for w in (16,):
    t0 = time.perf_counter(); ids = m.load_indexes(paths, workers=w); m.sync(); t1 = time.perf_counter()
    print(f"load_indexes workers={w}: {t1-t0:.2f} s = {total/(t1-t0)/1e9:.2f} GB/s", flush=True)
    assert m.download_index(ids[0])[:64] == chk
    for i in ids: m.evict(i)
raw = [p[:-3] for p in paths]
for w in (1, 4, 8):
    t0 = time.perf_counter(); ids = m.load_indexes(raw, workers=w); m.sync(); t1 = time.perf_counter()
    print(f"uncompressed files, load_indexes workers={w}: {t1-t0:.2f} s = {total/(t1-t0)/1e9:.2f} GB/s", flush=True)
    assert m.download_index(ids[0])[:64] == chk
    for i in ids: m.evict(i)
