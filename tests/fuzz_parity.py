#!/usr/bin/env python3
"""Randomised parity run against the CPU oracle (run on the GPU box):
random document counts / signature sizes / hash counts / k / canonical flag, query lengths across
every kernel class boundary, random thresholds and top-N.  Usage: fuzz_parity.py [cases] [seed]"""
import os, random, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import numpy as np
import oracle
from phylign_b200.matcher import Matcher

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rnd = random.Random(seed)
m = Matcher(0)
td = tempfile.mkdtemp()
checked_units = checked_hits = 0
for case in range(n_cases):
    n_docs = rnd.choice([1, 3, 8, 9, 64, 127, 128, 129, 250, 511, 513, 1000, 1024, 1025, 2000, 3999, 4000, 4096,
                         4097, 6000, rnd.randrange(1, 9000)])
    k = rnd.choice([31, 31, 31, 21, 15, rnd.randrange(8, 32)])
    nh = rnd.choice([1, 1, 1, 2, 3])
    canon = rnd.choice([1, 1, 0])
    sig = rnd.choice([1, 2, 61, 256, 1021, 4099, 65536, rnd.randrange(50, 20000)])
    glen = rnd.choice([400, 1500, 3500])
    root = "".join(rnd.choice("ACGT") for _ in range(glen))
    docs = []
    n_real = rnd.randrange(1, 12)
    for d in range(n_docs):
        if d < n_real or rnd.random() < 4.0 / n_docs:
            s = list(root)
            for _ in range(rnd.randrange(0, glen // 20)):
                s[rnd.randrange(glen)] = rnd.choice("ACGT")
            docs.append("".join(s).encode())
        else:
            docs.append(b"")
    oi = oracle.OracleIndex.construct(docs, term_size=k, canonicalize=canon, num_hashes=nh, signature_size_override=sig)
    body = oi.body
    dens = rnd.choice([0, 2, 3])
    if dens:
        noise = np.random.default_rng(case).integers(0, 256, size=body.shape, dtype=np.uint8)
        for _ in range(dens - 1):
            noise &= np.random.default_rng(case * 7 + _).integers(0, 256, size=body.shape, dtype=np.uint8)
        body |= noise
        if n_docs % 8:
            body[:, -1] &= (1 << (n_docs % 8)) - 1
    p = os.path.join(td, "f.cobs_classic")
    oi.write(p)
    for i in list(m.indexes):
        m.evict(i)
    idx = m.load_index(p, batch="fuzz__01")
    records = []
    for j in range(rnd.randrange(1, 14)):
        ln = rnd.choice([k - 1, k, k + 7, 100, 150, 254 + k, 255 + k, 256 + k, 600, 1022 + k, 1023 + k, 1024 + k,
                         glen, rnd.randrange(1, glen + 1)])
        ln = min(ln, glen)
        a = rnd.randrange(0, glen - ln + 1)
        records.append((f"q{j}", root[a:a + ln]))
    m.set_queries(records)
    thr = rnd.choice([0.0, 0.1, 0.33, 0.5, 0.7, 0.7, 0.9, 1.0, rnd.random()])
    top_n = rnd.choice([0, 1, 2, 5, 100])
    fl = rnd.random() < 0.2
    res = m.match(thr, top_n, fl)
    units = {int(u["query"]): u for u in res.units_of(idx)}
    for q, (_, s) in enumerate(records):
        if len(s) < k:
            assert q not in units, (case, q)
            continue
        kk, hits = oi.query(s.encode(), thr, fl)
        n_pass = len(hits)
        if top_n and n_pass > top_n:
            cut = hits[top_n - 1][1]
            hits = [h for h in hits if h[1] >= cut]
        if n_pass == 0:
            assert q not in units, (case, q, n_docs, k, nh, sig, thr)
            continue
        u = units[q]
        got = [(int(h["doc"]), int(h["score"])) for h in res.hits_of(u)]
        assert int(u["n_pass"]) == n_pass and got == hits, (case, q, n_docs, k, nh, canon, sig, len(s), thr, top_n, fl)
        checked_units += 1
        checked_hits += len(hits)
    if case % 25 == 24:
        print(f"case {case + 1}: ok so far ({checked_units} units, {checked_hits} hits compared)", flush=True)
print(f"fuzz parity: {n_cases} random cases (seed {seed}) bit-exact against the oracle; "
      f"{checked_units} non-empty units, {checked_hits} hits compared")
