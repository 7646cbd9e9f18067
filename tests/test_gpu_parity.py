"""Parity of the CUDA path against the CPU oracle and the golden vectors (needs a B200).

Everything goes through the C ABI (ctypes -> libphylign_cuda.so).  Bit-exact bar: integer
scores, hit sets, orders, header counts, merged candidate lists.
"""
import os
import random

import numpy as np
import pytest

import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    from phylign_b200.matcher import Matcher
    m = Matcher(0)
    yield m
    m.close()


@pytest.fixture(scope="module")
def golden_queries():
    return H.read_fasta(os.path.join(H.GOLDEN, "queries.fa"))


def _evict_all(m):
    for i in list(m.indexes):
        m.evict(i)


def _oracle_unit(oidx, seq, threshold, top_n, floor_mode=False):
    """(n_pass, [(doc, score)] kept) the way cobs + postprocess_cobs.py produce them."""
    k, hits = oidx.query(seq, threshold, floor_mode)
    n_pass = len(hits)
    if top_n and len(hits) > top_n:
        cut = hits[top_n - 1][1]
        hits = [h for h in hits if h[1] >= cut]
    return n_pass, hits


def _check_against_oracle(m, idx_id, oidx, records, threshold, top_n, floor_mode=False):
    res = m.match(threshold, top_n, floor_mode)
    units = {int(u["query"]): u for u in res.units_of(idx_id)}
    for q, (_, seq) in enumerate(records):
        seq = seq if isinstance(seq, bytes) else seq.encode()
        if len(seq) < oidx.term_size:
            assert q not in units
            continue
        n_pass, want = _oracle_unit(oidx, seq, threshold, top_n, floor_mode)
        if n_pass == 0:
            assert q not in units, f"query {q}: spurious unit"
            continue
        u = units[q]
        assert int(u["n_pass"]) == n_pass, f"query {q}"
        got = [(int(h["doc"]), int(h["score"])) for h in res.hits_of(u)]
        assert got == want, f"query {q}"
    return res


# --------------------------------------------------------------------------- golden vectors
def test_golden_cobs_text_and_match_files(M, golden_queries):
    from text_twins import format_cobs_text
    _evict_all(M)
    ids = {b: M.load_index(os.path.join(H.GOLDEN, f"{b}.cobs_classic.xz")) for b in H.GOLDEN_BATCHES}
    M.set_queries(golden_queries)
    res = M.match(0.7, top_n=0)
    for b, i in ids.items():
        assert format_cobs_text(golden_queries, res, M.indexes[i]) == H.golden_cobs_text(b)
    for keep in (1, 3, 100):
        res = M.match(0.7, top_n=keep)
        for b, i in ids.items():
            got = format_cobs_text(golden_queries, res, M.indexes[i], strip_prefix=True)
            assert got == H.golden_match_text(b, keep), (b, keep)


def test_golden_filter_fasta(M, golden_queries):
    """bit-exact intermediate/04_filter content vs the unmodified filter_queries.py."""
    from phylign_b200.cobs_index import ref_of
    from text_twins import format_filter_fasta
    _evict_all(M)
    # load in non-sorted order: the merge must not depend on it (filter_queries.py:135)
    for b in reversed(H.GOLDEN_BATCHES):
        M.load_index(os.path.join(H.GOLDEN, f"{b}.cobs_classic.xz"))
    names = M.set_ranks()
    refs = {ix.batch_rank: [ref_of(n) for n in ix.doc_names] for ix in M.indexes.values()}
    assert names == sorted(H.GOLDEN_BATCHES)
    M.set_queries(golden_queries)
    recs = [(h.split(" ")[0], s) for h, s in golden_queries]
    for keep in (1, 3, 100):
        M.match_run(0.7, top_n=keep, merge_top_n=keep)
        offs, cands = M.merged()
        assert format_filter_fasta(recs, offs, cands, refs) == H.golden_filter_fa(keep), keep


def test_dense_scores_bit_exact(M, golden_queries):
    _evict_all(M)
    for b in H.GOLDEN_BATCHES:
        raw = H.golden_index_bytes(b)
        i = M.load_index_bytes(raw, b)
        oidx = oracle.OracleIndex.parse(raw)
        M.set_queries(golden_queries)
        got = M.scores(i)
        for q, (_, s) in enumerate(golden_queries):
            k, want = oidx.scores(s.encode(), sliced=True)
            assert (got[q] == want).all(), (b, q)
        M.evict(i)


# --------------------------------------------------------------------------- shapes / edge cases
SHAPES = [  # n_docs, num_hashes, canonicalize, term_size
    (1, 1, 1, 31), (7, 1, 1, 31), (8, 2, 1, 31), (100, 1, 0, 31), (129, 3, 1, 31),
    (257, 1, 1, 21), (600, 1, 1, 31), (1000, 2, 1, 31), (2049, 1, 1, 31), (4000, 1, 1, 31),
    (4100, 1, 1, 31), (9001, 2, 1, 15),
]


@pytest.mark.parametrize("n_docs,nh,canon,k", SHAPES)
def test_random_indexes_vs_oracle(M, n_docs, nh, canon, k, tmp_path):
    rnd = random.Random(n_docs * 7 + nh)
    root = "".join(rnd.choice("ACGT") for _ in range(700))
    n_real = min(n_docs, 24)       # documents with sequence; the rest stay empty columns
    docs = []
    for d in range(n_docs):
        if d % max(1, n_docs // n_real) == 0 and len([x for x in docs if x]) < n_real:
            s = list(root)
            for _ in range(rnd.randrange(0, 60)):
                s[rnd.randrange(len(s))] = rnd.choice("ACGT")
            docs.append("".join(s).encode())
        else:
            docs.append(b"")
    sig = rnd.choice([97, 128, 1021, 4096])
    oidx = oracle.OracleIndex.construct(docs, term_size=k, canonicalize=canon, num_hashes=nh,
                                        signature_size_override=sig)
    # sprinkle noise so unrelated columns are not empty (false-positive-like bits)
    body = oidx.body
    noise = np.random.default_rng(n_docs).integers(0, 256, size=body.shape, dtype=np.uint8)
    noise &= np.random.default_rng(n_docs + 1).integers(0, 256, size=body.shape, dtype=np.uint8)
    body |= noise
    if n_docs % 8:
        body[:, -1] &= (1 << (n_docs % 8)) - 1
    p = tmp_path / "i.cobs_classic"
    oidx.write(p)
    _evict_all(M)
    i = M.load_index(str(p), batch="rnd__01")
    assert M.download_index(i) == body.tobytes()
    records = [("q%d" % j, root[a:a + ln]) for j, (a, ln) in enumerate(
        [(0, 150), (10, k), (20, k + 1), (5, 700 - 5), (300, 64), (100, k - 1), (50, 200)])]
    records.append(("rc", H.revcomp(root[40:240])))
    records.append(("noise", "".join(rnd.choice("ACGT") for _ in range(120))))
    M.set_queries(records)
    got = M.scores(i)
    for q, (_, s) in enumerate(records):
        kk, want = oidx.scores(s.encode(), sliced=True)
        assert (got[q] == want).all(), q
    for thr, top_n, fl in [(0.7, 0, False), (0.3, 5, False), (0.0, 3, False), (1.0, 1, False),
                           (0.55, 2, True)]:
        _check_against_oracle(M, i, oidx, records, thr, top_n, fl)


def test_long_queries_and_many_queries(M):
    """K > 1023 takes the general path; mixed lengths share one launch."""
    spec = oracle.SynthSpec(seed=77, n_docs=300, genome_len=6000, clade_size=10, clade_sub_q16=655,
                            doc_sub_q16=655)
    docs = [oracle.synth_genome(spec, d) for d in range(spec.n_docs)]
    oidx = oracle.OracleIndex.construct(docs)
    _evict_all(M)
    raw_path = "/tmp/phy_long.cobs_classic"
    oidx.write(raw_path)
    i = M.load_index(raw_path, batch="long__01")
    rnd = random.Random(4)
    records = []
    for j in range(200):
        d = rnd.randrange(spec.n_docs)
        ln = rnd.choice([31, 32, 100, 150, 500, 1053, 1054, 1500, 3000, 5999])
        a = rnd.randrange(0, 6000 - ln + 1)
        s = docs[d][a:a + ln].decode()
        records.append((f"r{j}", s if j % 2 else H.revcomp(s)))
    M.set_queries(records)
    _check_against_oracle(M, i, oidx, records, 0.7, 10)
    _check_against_oracle(M, i, oidx, records, 0.4, 0)
    os.unlink(raw_path)


def test_errors_fail_loudly(M):
    from phylign_b200._lib import PhylignCudaError
    _evict_all(M)
    oidx = oracle.OracleIndex.construct([b"ACGT" * 20, b"GGGTTTAAACCC" * 8])
    oidx.write("/tmp/phy_err.cobs_classic")
    raw = open("/tmp/phy_err.cobs_classic", "rb").read()
    with pytest.raises(Exception):
        M.load_index_bytes(raw[:-3], "bad__01")            # body too short
    assert not M.indexes
    with pytest.raises(PhylignCudaError) as e:
        M.set_queries([("q", "ACGT" * 10)])
        M.match(0.7)
    assert "PHY_ERR_STATE" in str(e.value)                  # no index resident
    i = M.load_index_bytes(raw, "ok__01")
    M.set_queries([("good", "ACGT" * 10), ("bad", "ACGTNACGT" * 5)])
    with pytest.raises(PhylignCudaError) as e:
        M.match(0.7)
    assert e.value.code == -6 and "#1" in str(e.value)      # cobs aborts on non-ACGT too
    M.set_queries([("lower", "acgt" * 10)])
    with pytest.raises(PhylignCudaError):
        M.match(0.7)
    M.set_queries([])
    res = M.match(0.7)
    assert len(res.units) == 0
    M.evict(i)
    os.unlink("/tmp/phy_err.cobs_classic")


# --------------------------------------------------------------------------- synthetic builder
def test_synth_index_and_reads_equal_oracle(M):
    from phylign_b200 import _lib
    _evict_all(M)
    kw = dict(seed=123, n_docs=150, genome_len=900, clade_size=16, clade_sub_q16=400, doc_sub_q16=300)
    ospec, dspec = oracle.SynthSpec(**kw), _lib.SynthSpec(**kw)
    docs = [oracle.synth_genome(ospec, d) for d in range(kw["n_docs"])]
    oidx = oracle.OracleIndex.construct(docs)
    i = M.add_synth_index("syn__01", dspec, oidx.signature_size)
    assert M.download_index(i) == oidx.body.tobytes()
    reads = M.synth_reads([dspec], 9, 5, 64, 150, random_q8=51, err_q16=655)
    for r in range(64):
        assert reads[r * 150:(r + 1) * 150] == oracle.synth_read([ospec], 9, 5 + r, 150, 51, 655), r
    records = [(f"s{r}", reads[r * 150:(r + 1) * 150].decode()) for r in range(64)]
    M.set_queries(records)
    _check_against_oracle(M, i, oidx, records, 0.7, 100)


def test_indexes_with_different_hash_counts_resident_together(M, tmp_path):
    """One match pass over indexes built with 1, 2 and 3 hash functions (same row width class): each is
    launched with its own AND depth, every unit exact, merged list == closed form over the oracle hits."""
    _evict_all(M)
    rnd = random.Random(77)
    spec = oracle.SynthSpec(seed=31, n_docs=900, genome_len=1500, clade_size=8, clade_sub_q16=500, doc_sub_q16=500)
    docs = [oracle.synth_genome(spec, d) for d in range(spec.n_docs)]
    oidxs, ids = {}, {}
    for nh in (1, 2, 3):
        o = oracle.OracleIndex.construct(docs, num_hashes=nh, doc_names=[f"{d:05d}_H{nh}D{d:05d}" for d in range(len(docs))])
        p = os.path.join(tmp_path, f"h{nh}__01.cobs_classic")
        o.write(p)
        oidxs[nh] = o
        ids[nh] = M.load_index(p)
    records = [(f"r{r}", docs[rnd.randrange(len(docs))][rnd.randrange(0, 300):][:rnd.choice([60, 150, 400, 1200])].decode())
               for r in range(48)]
    M.set_queries(records)
    for thr, top_n in ((0.7, 5), (0.3, 0)):
        res = M.match(thr, top_n)
        for nh in (1, 2, 3):
            units = {int(u["query"]): u for u in res.units_of(ids[nh])}
            for q, (_, seq) in enumerate(records):
                n_pass, want = _oracle_unit(oidxs[nh], seq.encode(), thr, top_n)
                if n_pass == 0:
                    assert q not in units
                    continue
                u = units[q]
                assert int(u["n_pass"]) == n_pass
                assert [(int(h["doc"]), int(h["score"])) for h in res.hits_of(u)] == want, (nh, q)
    _evict_all(M)


# --------------------------------------------------------------------------- size-independent properties
def test_properties_at_scale(M):
    """Config-2-like size (4000 docs): properties that need no oracle run.

    * a read copied from document d scores K in column d (no false negatives in a Bloom filter)
    * reverse complement of a query gives identical results (canonical k-mers)
    * concatenating two queries' k-mer sets: score(q1+q2 parts) additivity on disjoint chunks
    * top-N + ties is a prefix-closed subset of the full result and n_pass is unchanged
    """
    from phylign_b200 import _lib
    _evict_all(M)
    spec = _lib.SynthSpec(seed=2, n_docs=4000, genome_len=20000, clade_size=32, clade_sub_q16=328,
                          doc_sub_q16=328)
    ospec = oracle.SynthSpec(seed=2, n_docs=4000, genome_len=20000, clade_size=32, clade_sub_q16=328,
                             doc_sub_q16=328)
    sig = oracle.signature_size(20000 - 30)
    i = M.add_synth_index("big__01", spec, sig)
    rnd = random.Random(8)
    picks = [(rnd.randrange(4000), rnd.randrange(0, 20000 - 1000)) for _ in range(64)]
    seqs = [oracle.synth_genome(ospec, d)[a:a + 1000].decode() for d, a in picks]
    M.set_queries([(f"q{j}", s) for j, s in enumerate(seqs)] +
                  [(f"rc{j}", H.revcomp(s)) for j, s in enumerate(seqs)])
    sc = M.scores(i)
    for j, (d, _) in enumerate(picks):
        assert sc[j, d] == 970
        assert (sc[j] == sc[64 + j]).all()
    # additivity over a split with k-1 overlap: kmers(s) = kmers(s[:530]) + kmers(s[500:])
    M.set_queries([("a", seqs[0][:530]), ("b", seqs[0][500:]), ("ab", seqs[0])])
    s3 = M.scores(i)
    assert (s3[0] + s3[1] == s3[2]).all()
    M.set_queries([(f"q{j}", s) for j, s in enumerate(seqs)])
    full = M.match(0.7, 0)
    top = M.match(0.7, 100)
    for uf, ut in zip(full.units, top.units):
        assert uf["query"] == ut["query"] and uf["n_pass"] == ut["n_pass"]
        hf, ht = full.hits_of(uf), top.hits_of(ut)
        assert (np.diff(hf["score"].astype(np.int64)) <= 0).all()
        n = len(ht)
        assert (hf[:n] == ht).all()
        if len(hf) > 100:
            assert n >= 100 and ht["score"][-1] == hf["score"][99]
            assert n == len(hf) or hf["score"][n] < ht["score"][-1]
        else:
            assert n == len(hf)


def test_result_buffers_regrow_and_rerun(golden_queries):
    """Undersized hit/unit buffers: the gather pass reports the needed size and is rerun."""
    import subprocess, sys
    code = (
        "import os, sys; sys.path.insert(0, %r)\n"
        "from phylign_b200.matcher import Matcher\n"
        "from tests.text_twins import format_cobs_text\n"
        "from tests import helpers as H\n"
        "m = Matcher(0); i = m.load_index(os.path.join(H.GOLDEN, 'aaa__01.cobs_classic.xz'))\n"
        "qs = H.read_fasta(os.path.join(H.GOLDEN, 'queries.fa')); m.set_queries(qs)\n"
        "res = m.match(0.7, 0)\n"
        "assert format_cobs_text(qs, res, m.indexes[i]) == H.golden_cobs_text('aaa__01')\n"
        "print('ok', len(res.hits))\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       env=dict(os.environ, PHY_TEST_TINY_CAPS="1"))
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stderr[-2000:]


def test_deterministic_output(M, golden_queries):
    """Same inputs -> byte-identical result arrays (units ordered on the device, hits sorted)."""
    _evict_all(M)
    for b in H.GOLDEN_BATCHES:
        M.load_index(os.path.join(H.GOLDEN, f"{b}.cobs_classic.xz"))
    M.set_ranks()
    M.set_queries(golden_queries * 20)
    runs = []
    for _ in range(3):
        M.match_run(0.5, top_n=5, merge_top_n=5)
        res = M.fetch()
        offs, cands = M.merged()
        runs.append((res.units.tobytes(), res.hits.tobytes(), offs.tobytes(), cands.tobytes()))
    # offsets inside `hits` depend on atomics (allocation order); compare content per unit instead
    def canon(r):
        units = np.frombuffer(r[0], dtype=res.units.dtype)
        hits = np.frombuffer(r[1], dtype=res.hits.dtype)
        return [(int(u["query"]), int(u["index"]), int(u["n_pass"]),
                 hits[int(u["offset"]):int(u["offset"]) + int(u["n_kept"])].tobytes()) for u in units]
    assert canon(runs[0]) == canon(runs[1]) == canon(runs[2])
    assert runs[0][2:] == runs[1][2:] == runs[2][2:]


def test_c_abi_state_and_argument_errors(M):
    """Misuse of the C ABI returns a status and a message, never crashes or guesses."""
    import ctypes as C
    from phylign_b200 import _lib
    L = _lib.load()
    _evict_all(M)
    ctx = M._ctx
    rp = C.POINTER(_lib.Results)()
    mp = C.POINTER(_lib.Merged)()
    idx = C.c_int()
    assert L.phy_index_begin(ctx, b"x__01", 32, 1, 100, 1, 8, C.byref(idx)) == -2          # k > 31
    assert L.phy_index_begin(ctx, b"x__01", 31, 1, 0, 1, 8, C.byref(idx)) == -2           # empty signature
    assert L.phy_index_begin(ctx, b"x__01", 31, 1, 1 << 32, 1, 8, C.byref(idx)) == -2     # >= 2^32 rows
    assert L.phy_index_begin(ctx, b"x__01", 31, 1, 100, 1, (1 << 20) + 1, C.byref(idx)) == -2
    assert L.phy_index_begin(ctx, b"x__01", 31, 1, 100, 1, 8, C.byref(idx)) == 0
    i = idx.value
    buf = (C.c_char * 100)()
    assert L.phy_index_commit(ctx, i) == -4                                               # nothing pushed yet
    assert b"expected" in L.phy_last_error(ctx)
    assert L.phy_index_push(ctx, i, buf, 60) == 0 and L.phy_index_push(ctx, i, buf, 41) == -4   # 101 > 100 bytes
    assert L.phy_index_push(ctx, i, buf, 40) == 0 and L.phy_index_commit(ctx, i) == 0
    assert L.phy_index_push(ctx, i, buf, 1) == -4                                         # after commit
    assert L.phy_index_set_ranks(ctx, i, 4096, None) == -2
    assert L.phy_index_commit(ctx, 99) == -2 and L.phy_index_evict(ctx, -1) == -2
    M.set_queries([("q", "ACGT" * 10)])
    assert L.phy_results_fetch(ctx, C.byref(rp)) == -4                                    # no match run yet
    p = _lib.MatchParams(-0.1, 0, 0)
    assert L.phy_match_run(ctx, C.byref(p), 0) == -2                                      # negative threshold
    p = _lib.MatchParams(0.7, 0, 0)
    assert L.phy_match_run(ctx, C.byref(p), 0) == 0
    assert L.phy_merged_fetch(ctx, C.byref(mp)) == -4                                     # merge was not requested
    assert L.phy_results_fetch(ctx, C.byref(rp)) == 0 and rp.contents.n_units == 0
    L.phy_results_free(rp)
    assert L.phy_ctx_set_option(ctx, b"nonsense", 1) == -2
    assert L.phy_index_evict(ctx, i) == 0 and L.phy_index_evict(ctx, i) == -2
    assert L.phy_ctx_create(C.byref(C.c_void_p()), 4096, 0) == -2                         # no such device


@pytest.mark.parametrize("nh", [1, 2, 3, 5])
def test_every_query_length_class_for_every_hash_count(nh):
    """K = 40 .. 20000 (8-, 10-, 14-plane and chunked classes) against indexes with 1, 2, 3 and 5 hash
    functions (the ring kernel ANDs the h rows of a k-mer as they leave the ring): no query may be
    dropped, every score exact."""
    import subprocess, sys
    code = r'''
import os, sys, random
sys.path.insert(0, %r)
import oracle
from phylign_b200.matcher import Matcher
from tests.test_gpu_parity import _check_against_oracle
nh = %d
spec = oracle.SynthSpec(seed=91, n_docs=700, genome_len=21000, clade_size=8, clade_sub_q16=655, doc_sub_q16=655)
docs = [oracle.synth_genome(spec, d) if d %% 50 == 0 else b"" for d in range(spec.n_docs)]
oidx = oracle.OracleIndex.construct(docs, num_hashes=nh, signature_size_override=30011)
oidx.write("/tmp/phy_cls.cobs_classic")
m = Matcher(0)
i = m.load_index("/tmp/phy_cls.cobs_classic", batch="cls__01")
g = docs[0].decode()
records = [(f"L{ln}", g[7:7 + ln]) for ln in (40, 70, 285, 286, 600, 1053, 1054, 5000, 16413, 16414, 20000)]
m.set_queries(records)
_check_against_oracle(m, i, oidx, records, 0.7, 3)
_check_against_oracle(m, i, oidx, records, 0.2, 0)
print("ok")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), nh)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout + r.stderr)[-3000:]
