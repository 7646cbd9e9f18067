"""Pin the CPU oracle: XXH64 KATs, brute force, layout, golden vectors (CPU only)."""
import json
import os
import random
import struct

import numpy as np
import pytest

import oracle
from oracle import filters
from tests import helpers as H


def test_xxh64_golden_kats(golden_dir):
    vecs = json.load(open(os.path.join(golden_dir, "xxh64_kat.json")))
    assert len(vecs) > 200
    for v in vecs:
        data = v["data"].encode()
        assert oracle.xxh64(data, v["seed"]) == int(v["hash"], 16)
        assert H.py_xxh64(data, v["seed"]) == int(v["hash"], 16)


def test_xxh64_survey_appendix_b():
    # SURVEY.md Appendix B (canonical form hashed, ASCII, len 31)
    kats = [("A" * 31, 0x04d4645ec33f5384, 0x04caeab334eeb221),
            ("T" * 31, 0x04d4645ec33f5384, 0x04caeab334eeb221),
            ("ACGTACGTACGTACGTACGTACGTACGTACG", 0x2e0e4ebd5477cb86, 0x12ade2aa4f60c39e),
            ("ATTGGTTCATGTAATAAGCCCTCATCAGAAT", 0xe8253be3ff32cf92, 0x27f106d72baf068f),
            ("TTGGTTCATGTAATAAGCCCTCATCAGAATG", 0xcbcd0d6fa840fcfa, 0xf7a1b6f914dd1020),
            ("TGAGGCGATCACCTGGTTGAACTGCTGCCGG", 0xd116b0400928e46e, 0x7341b680c6eddb05)]
    mods = [2000848, 2000848, 8852274, 13697892, 9971052, 12900968]
    for (kmer, h0, h1), m in zip(kats, mods):
        c = oracle.canonical(kmer.encode())
        assert c.decode() == min(kmer, H.revcomp(kmer))
        assert oracle.xxh64(c, 0) == h0
        assert oracle.xxh64(c, 1) == h1
        assert h0 % 21188834 == m


def test_xxh64_random_vs_python_xxhash():
    xxhash = pytest.importorskip("xxhash")
    rnd = random.Random(3)
    for _ in range(500):
        n = rnd.randrange(0, 200)
        data = bytes(rnd.randrange(256) for _ in range(n))
        seed = rnd.randrange(1 << 64)
        assert oracle.xxh64(data, seed) == xxhash.xxh64(data, seed=seed).intdigest()


def test_canonical_and_invalid():
    rnd = random.Random(5)
    for _ in range(300):
        k = rnd.choice([1, 2, 15, 16, 31, 32, 33])
        s = "".join(rnd.choice("ACGT") for _ in range(k))
        assert oracle.canonical(s.encode()).decode() == min(s, H.revcomp(s))
    assert oracle.canonical(b"ACGTN" + b"A" * 26) is None
    assert oracle.canonical(b"acgt" + b"A" * 27) is None


def test_threshold_examples():
    # SURVEY.md A.4/A.6 worked examples in IEEE double
    assert oracle.threshold_terms(0.7, 120) == 84
    assert oracle.threshold_terms(0.7, 970) == 679
    assert oracle.threshold_terms(0.7, 121) == 85
    assert oracle.threshold_terms(0.7, 121, floor_mode=True) == 84
    assert oracle.threshold_terms(0.0, 100) == 0
    assert oracle.threshold_terms(1.0, 100) == 100


def test_signature_size():
    assert oracle.signature_size(1_000_000, 1, 0.3) == int(np.ceil(1_000_000 * (-1 / np.log(1 - 0.3))))
    assert abs(oracle.signature_size(10 ** 6, 1, 0.3) / 1e6 - 2.8037) < 1e-3


@pytest.mark.parametrize("num_hashes,canon", [(1, 1), (2, 1), (3, 0)])
def test_scores_vs_brute_force(num_hashes, canon):
    rnd = random.Random(10 + num_hashes)
    docs = []
    root = "".join(rnd.choice("ACGT") for _ in range(300))
    for d in range(21):
        s = list(root)
        for _ in range(d):
            s[rnd.randrange(len(s))] = rnd.choice("ACGT")
        docs.append("".join(s))
    sig = 701
    idx = oracle.OracleIndex.construct([d.encode() for d in docs], num_hashes=num_hashes,
                                       canonicalize=canon, signature_size_override=sig)
    for q in (root[10:160], H.revcomp(root[50:200]), docs[20][:100], root[:31]):
        want = H.brute_scores(docs, q, num_hashes=num_hashes, sig=sig, canonicalize=bool(canon))
        k, got = idx.scores(q.encode())
        assert k == len(q) - 30
        assert got.tolist() == want
        k2, got2 = idx.scores(q.encode(), sliced=True, threads=3)
        assert k2 == k and got2.tolist() == want


def test_sliced_equals_scalar_many_docs():
    rnd = random.Random(77)
    docs = ["".join(rnd.choice("ACGT") for _ in range(400)) for _ in range(300)]
    idx = oracle.OracleIndex.construct([d.encode() for d in docs])
    assert idx.row_size == 38
    for d in (0, 127, 128, 299):
        q = docs[d][17:217].encode()
        k, a = idx.scores(q)
        _, b = idx.scores(q, sliced=True, threads=4)
        assert (a == b).all() and a[d] == k


def test_short_invalid_and_empty_queries():
    idx = oracle.OracleIndex.construct([b"ACGT" * 30])
    assert idx.scores(b"ACGT" * 7)[0] == 0               # 28 < k
    assert idx.scores(b"")[0] == 0
    assert idx.scores(b"ACGTN" * 10)[0] == -1
    assert idx.query_text([("e", b""), ("s", b"ACGTACGT")], 0.7) == "*s\t0\n"


def test_header_layout_roundtrip(tmp_path):
    names = ["000001_SAMA", "000002_SAMB", "000777_SAMC"]
    idx = oracle.OracleIndex.construct([b"ACGT" * 20, b"GATTACA" * 12, b"TTTTGGGGCCCCAAAA" * 5], names)
    p = tmp_path / "x.cobs_classic"
    idx.write(p)
    raw = p.read_bytes()
    # SURVEY.md A.1
    assert raw[:18] == b"COBS:CLASSIC_INDEX"
    ver, k, canon, nd, sig, nh = struct.unpack_from("<IIBIQQ", raw, 18)
    assert (ver, k, canon, nd, nh) == (1, 31, 1, 3, 1) and sig == idx.signature_size
    p0 = 18 + 29
    blob = b"".join(n.encode() + b"\n" for n in names)
    assert raw[p0:p0 + len(blob)] == blob
    assert raw[p0 + len(blob):p0 + len(blob) + 13] == b"CLASSIC_INDEX"
    assert idx.header_size == p0 + len(blob) + 13
    assert len(raw) == idx.header_size + sig * 1
    back = oracle.OracleIndex.parse(raw)
    assert back.doc_names == names and (back.body == idx.body).all()
    with pytest.raises(ValueError):
        oracle.OracleIndex.parse(raw[:-1])
    # bit order: doc d <-> byte d/8, bit d%8 (LSB first)
    row = oracle.xxh64(oracle.canonical(b"ACGT" * 7 + b"ACG"), 0) % sig
    assert idx.body[row, 0] & 1


@pytest.mark.parametrize("batch", H.GOLDEN_BATCHES)
def test_golden_cobs_text_reproduced(batch):
    idx = oracle.OracleIndex.parse(H.golden_index_bytes(batch))
    recs = [(h, s.encode()) for h, s in H.read_fasta(os.path.join(H.GOLDEN, "queries.fa"))]
    assert idx.query_text(recs, 0.7) == H.golden_cobs_text(batch)


@pytest.mark.parametrize("keep", [1, 3, 100])
def test_filters_match_reference_scripts(keep):
    """oracle/filters.py == outputs of the unmodified reference scripts."""
    queries = [(h.split(" ")[0], s) for h, s in H.read_fasta(os.path.join(H.GOLDEN, "queries.fa"))]
    batches = []
    for batch in H.GOLDEN_BATCHES:
        post = filters.postprocess_text(H.golden_cobs_text(batch), keep)
        assert post == H.golden_match_text(batch, keep)
        per_query = []
        for head, _n, hits in filters.parse_cobs_text(post):
            per_query.append((head.split(" ")[0], [(nm.split("_")[1], sc) for nm, sc in hits]))
        batches.append((batch, per_query))
    assert filters.merge_running(queries, batches, keep) == H.golden_filter_fa(keep)
    assert filters.merge_closed_form(queries, batches, keep) == H.golden_filter_fa(keep)
    # argv order of the match files does not matter (filter_queries.py:135)
    assert filters.merge_running(queries, batches[::-1], keep) == H.golden_filter_fa(keep)


def test_merge_closed_form_equals_running_random():
    rnd = random.Random(1)
    for trial in range(200):
        keep = rnd.choice([1, 2, 3, 5, 10])
        queries = [(f"q{i}", "ACGT") for i in range(3)]
        batches = []
        for b in range(rnd.randrange(1, 6)):
            per_query = []
            for q, _ in queries:
                hits = sorted(((f"x_R{b}{rnd.randrange(50):02d}", rnd.randrange(1, 8))
                               for _ in range(rnd.randrange(0, 12))), key=lambda x: -x[1])
                hits = list({h[0]: h for h in hits}.values())
                hits.sort(key=lambda x: -x[1])
                kept = filters.postprocess_block(hits, keep)
                per_query.append((q, [(nm.split("_")[1], sc) for nm, sc in kept]))
            batches.append((f"b{rnd.randrange(100):02d}__{b:02d}", per_query))
        assert filters.merge_running(queries, batches, keep) == \
            filters.merge_closed_form(queries, batches, keep)


def test_synth_spec_is_deterministic():
    s = oracle.SynthSpec(seed=5, n_docs=40, genome_len=600, clade_size=8, clade_sub_q16=655,
                         doc_sub_q16=655)
    g0, g1, g9 = (oracle.synth_genome(s, d) for d in (0, 1, 9))
    assert g0 == oracle.synth_genome(s, 0) and len(g0) == 600
    diff01 = sum(a != b for a, b in zip(g0, g1))
    diff09 = sum(a != b for a, b in zip(g0, g9))
    assert 0 < diff01 < 40 and diff01 <= diff09 < 80      # same clade closer than other clade
    r = oracle.synth_read([s], 1, 3, 150, random_q8=0, err_q16=0)
    gs = [oracle.synth_genome(s, d).decode() for d in range(40)]
    assert any(r.decode() in g or H.revcomp(r.decode()) in g for g in gs)


def test_avx2_counting_equals_sse2_and_bit_loop():
    """The 256-bit counting kernel (CPU-baseline option) gives the scores of the SSE2 (cobs-shape) one."""
    import random
    rnd = random.Random(11)
    docs = [bytes(rnd.choice(b"ACGT") for _ in range(400)) for _ in range(300)]   # 300 docs: 38-B rows (ragged slices)
    idx = oracle.OracleIndex.construct(docs)
    reads = [docs[d][10:250] for d in (0, 17, 299)] + [bytes(rnd.choice(b"ACGT") for _ in range(120))]
    want = [idx.scores(r)[1] for r in reads]
    try:
        for avx in (False, True):
            mode = oracle.set_simd(avx)
            for r, w in zip(reads, want):
                for threads in (1, 3):
                    assert (idx.scores(r, sliced=True, threads=threads)[1] == w).all(), (mode, threads)
    finally:
        oracle.set_simd(False)


def test_oracle_cli_thread_layouts_print_the_same_text(tmp_path):
    """cobs_oracle query: queries-over-threads (default for -T > 1) == slices-over-threads == serial."""
    import subprocess
    idxp = os.path.join(tmp_path, "i.cobs_classic")
    open(idxp, "wb").write(H.golden_index_bytes("bbb__01"))
    q = os.path.join(H.GOLDEN, "queries.fa")
    outs = [subprocess.run([oracle.CLI_PATH, "query", "-t", "0.7", "-T", t, "-i", idxp, "-f", q] + extra,
                           capture_output=True, text=True, check=True).stdout
            for t, extra in (("1", []), ("4", []), ("4", ["--threads-over-slices"]), ("4", ["--avx2"]))]
    assert outs[0] == H.golden_cobs_text("bbb__01")
    assert outs[1] == outs[0] and outs[2] == outs[0] and outs[3] == outs[0]
