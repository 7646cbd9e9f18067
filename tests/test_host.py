"""CPU-only tests of the host side: C-ABI surface, index container, readers, text formats."""
import gzip
import os
import re

import numpy as np
import pytest

import oracle
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from phylign_b200 import _lib, build
    build.build()
    L = _lib.load()
    header = open(os.path.join(ROOT, "include", "phylign_cuda.h")).read()
    declared = set(re.findall(r"\b(phy_[a-z0-9_]+)\s*\(", header))
    declared -= {"phy_ctx"}
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/phylign_cuda.h but not exported"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype"
    assert L.phy_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU the product must fail loudly, never compute on the CPU."""
    import ctypes as C
    from phylign_b200 import _lib
    L = _lib.load()
    n = C.c_int(-1)
    rc = L.phy_device_count(C.byref(n))
    if rc == 0:
        pytest.skip("a GPU is visible here")
    assert rc == -1 and n.value == 0
    from phylign_b200.matcher import Matcher
    with pytest.raises(_lib.PhylignCudaError):
        Matcher(0)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "phylign_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, re.M), fn
                assert "liboracle" not in src and "cobs_oracle.h" not in src, fn


@pytest.mark.parametrize("batch", H.GOLDEN_BATCHES)
def test_index_stream_parses_golden_xz(batch):
    from phylign_b200.cobs_index import IndexStream, parse_bytes
    raw = H.golden_index_bytes(batch)
    oidx = oracle.OracleIndex.parse(raw)
    with IndexStream(os.path.join(H.GOLDEN, f"{batch}.cobs_classic.xz"), chunk_bytes=10007) as st:
        h = st.header
        body = b"".join(bytes(c) for c in st.body_chunks())
    assert (h.term_size, h.canonicalize, h.n_docs, h.signature_size, h.num_hashes) == \
        (31, 1, oidx.n_docs, oidx.signature_size, 1)
    assert h.doc_names == oidx.doc_names and h.header_size == oidx.header_size
    assert body == oidx.body.tobytes()
    h2, body2 = parse_bytes(raw)
    assert body2 == body and h2.to_bytes() == raw[:h.header_size]


def test_index_stream_rejects_bad_sizes(tmp_path):
    from phylign_b200.cobs_index import IndexFormatError, IndexStream
    raw = H.golden_index_bytes("bbb__01")
    for name, data in (("short", raw[:-5]), ("long", raw + b"xx"), ("magic", b"XOBS" + raw[4:])):
        p = tmp_path / f"{name}.cobs_classic"
        p.write_bytes(data)
        with pytest.raises(IndexFormatError):
            with IndexStream(str(p)) as st:
                for _ in st.body_chunks():
                    pass


def test_fasta_readers(tmp_path):
    from phylign_b200.fasta import read_cobs_records, read_fastx
    p = tmp_path / "q.fa"
    p.write_text(">a desc here\nACGT\nAC\n\n;b\nGG\n>empty\n>c\nTT\n")
    assert read_cobs_records(p) == [("a desc here", "ACGTAC"), ("b", "GG"), ("c", "TT")]
    assert read_cobs_records(os.path.join(H.GOLDEN, "queries.fa")) == \
        H.read_fasta(os.path.join(H.GOLDEN, "queries.fa"))
    fq = tmp_path / "r.fq"
    fq.write_text("@r1 x\nACGT\n+\nIIII\n@r2\nGG\nTT\n+r2\nII\nII\n")
    assert read_fastx(fq) == [("r1", "ACGT"), ("r2", "GGTT")]
    assert [n for n, _ in read_fastx(p)] == ["a", "empty", "c"]  # ';' is not a readfq header


def test_text_formatters_with_synthetic_results():
    from text_twins import format_cobs_text, format_filter_fasta
    from phylign_b200.matcher import CAND_DT, HIT_DT, UNIT_DT, MatchResult, ResidentIndex
    from phylign_b200.cobs_index import ClassicHeader
    hdr = ClassicHeader(31, 1, 3, 10, 1, ["zz9_SAMEA3", "ab1_SAMEA1", "qq2_SAMEA2"])
    ix = ResidentIndex(0, "bb__01", hdr)
    units = np.array([(0, 0, 3, 2, 0), (2, 0, 1, 1, 2)], dtype=UNIT_DT)
    hits = np.array([(0, 6), (1, 6), (2, 8)], dtype=HIT_DT)
    res = MatchResult(units, hits, np.array([10, 10, 10, 0], np.uint32), 4, 0, 0)
    recs = [("q1 c", "A" * 40), ("q2", "C" * 40), ("q3", "G" * 40), ("e", "")]
    assert format_cobs_text(recs, res, ix) == \
        "*q1 c\t3\nzz9_SAMEA3\t6\nab1_SAMEA1\t6\n*q2\t0\n*q3\t1\nqq2_SAMEA2\t8\n"
    assert format_cobs_text(recs, res, ix, strip_prefix=True) == \
        "*q1 c\t3\n_SAMEA3\t6\n_SAMEA1\t6\n*q2\t0\n*q3\t1\n_SAMEA2\t8\n"
    # SURVEY.md Appendix C vector
    cands = np.array([(6, 0, 0, 0), (6, 1, 1, 0), (6, 1, 0, 1), (8, 0, 1, 0), (8, 1, 2, 0)], dtype=CAND_DT)
    offs = np.array([0, 3, 3, 5], dtype=np.uint64)
    refs = {0: ["SAMEB3", "SAMEB5"], 1: ["SAMEA3", "SAMEA1", "SAMEA5"]}
    got = format_filter_fasta([("q1", "AC"), ("q2", "GT"), ("q3", "TT")], offs, cands, refs)
    assert got == ">q1 SAMEB3,SAMEA1,SAMEA3\nAC\n>q2 \nGT\n>q3 SAMEB5,SAMEA5\nTT\n"


@pytest.mark.parametrize("keep", [1, 3, 100])
def test_cli_postprocess_stream_equals_reference_output(keep):
    """`phylign_b200.cli postprocess -n N` == unmodified postprocess_cobs.py (golden files)."""
    import io
    from phylign_b200.cli import _postprocess_stream
    for batch in H.GOLDEN_BATCHES:
        out = io.StringIO()
        _postprocess_stream(io.StringIO(H.golden_cobs_text(batch)), out, keep)
        assert out.getvalue() == H.golden_match_text(batch, keep)


def test_cli_parse_match_file_rules(tmp_path):
    from text_twins import parse_match_file
    p = os.path.join(H.GOLDEN, "n3", "aaa__01____queries.gz")
    blocks = parse_match_file(p)
    ref = H.read_fasta(os.path.join(H.GOLDEN, "queries.fa"))
    assert [q for q, _ in blocks] == [h.split(" ")[0] for h, _ in ref]
    bad = tmp_path / "b__01____q.gz"
    with gzip.open(bad, "wt") as f:
        f.write("*q1\t1\nno_under_score_name\t5\n")
    with pytest.raises(ValueError):
        parse_match_file(str(bad))          # filter_queries.py:64 raises on a second underscore too
    empty = tmp_path / "e__01____q.gz"
    with gzip.open(empty, "wt") as f:
        f.write("")
    with pytest.raises(ValueError):
        parse_match_file(str(empty))


def test_run_cobs_streaming_usage_error():
    import subprocess
    r = subprocess.run([os.path.join(ROOT, "scripts", "run_cobs_streaming.sh"), "0.7", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "usage:" in r.stderr and "kmer_thres threads cobs_index.xz" in r.stderr


def test_fix_query_matches_snakefile_rule(tmp_path):
    """fix-query == `seqtk seq -A -U -C | awk gsub(/[^ACGT]/,"A")` (Snakefile:326-332)."""
    from phylign_b200.fasta import fix_query_file, fix_query_seq
    assert fix_query_seq("acgtNnRyACGT-*") == "ACGTAAAAACGTAA"
    fq = tmp_path / "r.fq"
    fq.write_text("@r1 some comment\nacgtnACGT\n+\nIIIIIIIII\n@r2\nGGNN\n+\nIIII\n")
    fa = tmp_path / "g.fa"
    fa.write_text(">g1 desc\nACGT\nryk\n>g2\nTTTT\n")
    assert fix_query_file(fq) == ">r1\nACGTAACGT\n>r2\nGGAA\n"
    assert fix_query_file(fa) == ">g1\nACGTAAA\n>g2\nTTTT\n"
    # the committed ARGannot fixture was produced with exactly this transform
    import lzma
    txt = lzma.open(os.path.join(H.GOLDEN, "ARGannot_r3.fixed.fa.xz"), "rt").read()
    assert set("".join(l for l in txt.splitlines() if not l.startswith(">"))) == set("ACGT")


def test_c_formatters_equal_python_formatters():
    """phy_format_cobs_text / phy_format_filter_fasta (C++, host only) == the pure-Python
    formatters, on fabricated result structs (no GPU involved)."""
    import ctypes as C
    from phylign_b200 import _lib
    from phylign_b200.cobs_index import ClassicHeader
    from phylign_b200.cobs_text import format_cobs_text_fast, format_filter_fasta_fast
    from text_twins import format_cobs_text, format_filter_fasta
    from phylign_b200.matcher import CAND_DT, HIT_DT, UNIT_DT, MatchResult, ResidentIndex
    import random
    rnd = random.Random(3)
    n_docs, nq = 37, 50
    names = [f"{rnd.randrange(10**6):06d}_ACC{d:04d}" for d in range(n_docs)] 
    names[5] = "nounderscore"          # remove_rnd_id on such a name yields "_" 
    ixs = [ResidentIndex(i, f"b__0{i}", ClassicHeader(31, 1, n_docs, 10, 1, names)) for i in (0, 2)]
    recs = [(f"q{q} comment {q}", "" if q % 11 == 0 else "ACGT" * 10) for q in range(nq)]
    units, hits = [], []
    for i in (0, 2):
        for q in range(nq):
            if rnd.random() < 0.5 and recs[q][1]:
                k = rnd.randrange(1, 9)
                units.append((q, i, k + rnd.randrange(3), k, len(hits)))
                hits += [(rnd.randrange(n_docs), rnd.randrange(1, 999999)) for _ in range(k)]
    ua, ha = np.array(units, dtype=UNIT_DT), np.array(hits, dtype=HIT_DT)
    nk = np.zeros(nq, np.uint32)
    res = MatchResult(ua, ha, nk, nq, 0, 0)
    r = _lib.Results(nq, 2, len(ua), C.cast(ua.ctypes.data, C.POINTER(_lib.Unit)), len(ha),
                     C.cast(ha.ctypes.data, C.POINTER(_lib.Hit)), C.cast(nk.ctypes.data, C.POINTER(C.c_uint32)), 0, 0)
    for ix in ixs:
        for strip in (False, True):
            want = format_cobs_text(recs, res, ix, strip_prefix=strip).encode()
            assert format_cobs_text_fast(recs, res, ix, strip_prefix=strip, results_ptr=C.pointer(r)) == want
    # merged
    refs = {0: [f"SAM{d}" for d in range(n_docs)], 1: [], 2: [f"ERR{d}" for d in range(n_docs)]}
    offs = np.zeros(nq + 1, np.uint64)
    cands = []
    for q in range(nq):
        for _ in range(rnd.randrange(0, 5)):
            b = rnd.choice([0, 2])
            cands.append((rnd.randrange(1, 500), b, rnd.randrange(n_docs), 0))
        offs[q + 1] = len(cands)
    ca = np.array(cands, dtype=CAND_DT)
    m = _lib.Merged(nq, C.cast(offs.ctypes.data, C.POINTER(C.c_uint64)), C.cast(ca.ctypes.data, C.POINTER(_lib.Cand)), 0)
    qrecs = [(f"q{q}", "ACGT" * (q % 5)) for q in range(nq)]
    assert format_filter_fasta_fast(qrecs, C.pointer(m), refs) == format_filter_fasta(qrecs, offs, ca, refs).encode()


def test_fast_modulo_algorithm_is_exact():
    """The multiply-high modulo of csrc/phy_internal.cuh (phy_fastmod), restated with Python
    integers: exact for every 64-bit hash and every signature_size < 2^32."""
    import random
    rnd = random.Random(9)
    M64 = (1 << 64) - 1

    def fastmod(n, d):
        magic = M64 // d                       # host side: UINT64_MAX / signature_size
        q = (n * magic) >> 64                  # __umul64hi
        r = (n - q * d) & M64
        if r >= d:
            r -= d
        return r

    ds = [1, 2, 3, 7, 97, 128, 4096, 65536, 1 << 31, (1 << 32) - 2, (1 << 32) - 5, 21188834, 2803644, 31_800_000]
    ds += [rnd.randrange(1, (1 << 32) - 1) for _ in range(300)]
    ns = [0, 1, M64, M64 - 1, 1 << 63, (1 << 63) - 1, 1 << 32, (1 << 32) - 1]
    for d in ds:
        for n in ns + [rnd.randrange(1 << 64) for _ in range(200)] + [d - 1, d, d + 1, 2 * d - 1, 2 * d, M64 // d * d]:
            n &= M64
            assert fastmod(n, d) == n % d, (n, d)


def test_candidate_buckets_equal_batch_align_load_qdicts():
    """The per-batch reference -> queries tables equal what batch_align.py:126-171 (load_qdicts)
    derives from the 04_filter FASTA (restated here on the golden file)."""
    from phylign_b200.cobs_index import parse_bytes, ref_of
    from phylign_b200.cobs_text import candidate_buckets, format_bucket_tsv
    from phylign_b200.matcher import CAND_DT
    keep = 3
    refs = {}
    for r, b in enumerate(sorted(H.GOLDEN_BATCHES)):
        hdr, _ = parse_bytes(H.golden_index_bytes(b))
        refs[r] = [ref_of(n) for n in hdr.doc_names]
    acc_to = {a: (r, d) for r, names in refs.items() for d, a in enumerate(names)}
    # parse the golden 04_filter FASTA the way readfq does: name, comment = candidate list
    qnames, rows, offs = [], [], [0]
    for line in H.golden_filter_fa(keep).splitlines():
        if line.startswith(">"):
            name, _, com = line[1:].partition(" ")
            qnames.append(name)
            for a in filter(len, com.split(",")):
                r, d = acc_to[a]
                rows.append((0, r, d, 0))
            offs.append(len(rows))
    cands = np.array(rows, dtype=CAND_DT)
    got = candidate_buckets(qnames, np.array(offs, dtype=np.uint64), cands, refs)
    # restatement of load_qdicts steps 2-3 per batch
    for r, names in refs.items():
        want = {}
        for line in H.golden_filter_fa(keep).splitlines():
            if line.startswith(">"):
                name, _, com = line[1:].partition(" ")
                for a in filter(len, com.split(",")):
                    if a in names:
                        want.setdefault(a, []).append(name)
        assert dict(got.get(r, [])) == want
        assert format_bucket_tsv(got.get(r, [])).count("\n") == len(want)


def test_header_roundtrip_property():
    """ClassicHeader.to_bytes / read_header round trip on random headers (hypothesis), including
    bodies whose first bytes look like a name table, and chunked/pipe-like reads."""
    import io
    from hypothesis import given, settings, strategies as st
    from phylign_b200.cobs_index import ClassicHeader, IndexFormatError, parse_bytes, read_header

    name = st.text(alphabet="ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_.-", min_size=1, max_size=24)

    class Dribble(io.RawIOBase):          # a pipe that hands out at most 7 bytes per read
        def __init__(self, b):
            self.b, self.p = b, 0

        def readable(self):
            return True

        def read(self, n=-1):
            n = 7 if n < 0 else min(n, 7)
            out = self.b[self.p:self.p + n]
            self.p += len(out)
            return out

    @settings(max_examples=60, deadline=None)
    @given(st.lists(name, min_size=1, max_size=40), st.integers(1, 31), st.integers(0, 1), st.integers(1, 50),
           st.integers(1, 4), st.binary(min_size=0, max_size=3))
    def check(names, k, canon, sig, nh, junk):
        hdr = ClassicHeader(k, canon, len(names), sig, nh, names)
        body = (b"\nCLASSIC_INDEX\n" * 50 + junk)[:hdr.body_size].ljust(hdr.body_size, b"\x55")
        raw = hdr.to_bytes() + body
        h2, b2 = parse_bytes(raw)
        assert (h2.term_size, h2.canonicalize, h2.n_docs, h2.signature_size, h2.num_hashes, h2.doc_names) == \
            (k, canon, len(names), sig, nh, names)
        assert b2 == body and h2.header_size == len(hdr.to_bytes())
        h3 = read_header(Dribble(raw))
        assert h3.doc_names == names and h3.header_size == h2.header_size
        try:
            parse_bytes(raw + b"x")
            assert False, "oversized body accepted"
        except IndexFormatError:
            pass

    check()


def test_native_match_file_parser_equals_python_parser(tmp_path):
    """phy_parse_match_text (C++) == the line-by-line Python restatement of filter_queries.py:27-66,
    on the golden match files and on malformed inputs (same inputs rejected)."""
    from phylign_b200.cli import parse_match_file_native
    from text_twins import parse_match_file
    for keep in (1, 3, 100):
        for b in H.GOLDEN_BATCHES:
            p = os.path.join(H.GOLDEN, f"n{keep}", f"{b}____queries.gz")
            want = parse_match_file(p)
            qnames, first_hit, ref_ids, refs_sorted, kmers = parse_match_file_native(p)
            assert [q.decode() for q in qnames.tolist()] == [q for q, _ in want]
            got = [[(refs_sorted[int(r)], int(k)) for r, k in zip(ref_ids[int(a):int(z)], kmers[int(a):int(z)])]
                   for a, z in zip(first_hit[:-1], first_hit[1:])]
            assert got == [hits for _, hits in want]
            assert refs_sorted == sorted(refs_sorted)
    cases = {"hit_first": "a_b\t3\n", "two_underscores": "*q\t1\na_b_c\t3\n", "no_underscore": "*q\t1\nabc\t3\n",
             "three_fields": "*q\t1\na_b\t3\t4\n", "no_tab_header": "*q 1\n", "bad_count": "*q\tx\n",
             "bad_kmers": "*q\t1\na_b\tx\n", "empty": "\n\n"}
    for name, text in cases.items():
        p = tmp_path / f"{name}__01____q.txt"
        p.write_text(text)
        with pytest.raises(Exception):
            parse_match_file(str(p))
        with pytest.raises(ValueError):
            parse_match_file_native(str(p))
    ok = tmp_path / "ok__01____q.txt"
    ok.write_text("\n*q1 some comment\t2  \n  x_R1 \t 5\r\ny_R0\t4\n\n*q2\t0\n")
    assert parse_match_file(str(ok)) == [("q1", [("R1", 5), ("R0", 4)]), ("q2", [])]
    qn, fh, ri, rs, km = parse_match_file_native(str(ok))
    assert qn.tolist() == [b"q1", b"q2"] and fh.tolist() == [0, 2, 2] and [rs[i] for i in ri] == ["R1", "R0"] and km.tolist() == [5, 4]


def test_query_name_index_vectorised_lookup():
    """Names of match-file blocks -> positions in the query file: in-order fast path, any order, unknown names."""
    from phylign_b200.cli import _QueryNameIndex
    qid = {"q1": 0, "longer_name_7": 1, "b": 2}
    ix = _QueryNameIndex(qid)
    assert ix.lookup(np.array([b"q1", b"longer_name_7", b"b"]), "x").tolist() == [0, 1, 2]
    assert ix.lookup(np.array([b"b", b"q1", b"b"], dtype="S2"), "x").tolist() == [2, 0, 2]
    assert ix.lookup(np.zeros(0, dtype="S1"), "x").tolist() == []
    with pytest.raises(KeyError):
        ix.lookup(np.array([b"q1", b"nope"]), "batch zz__01")
    with pytest.raises(KeyError):
        ix.lookup(np.array([b"q1", b"longer_name_7_and_more", b"b"]), "x")
    assert _QueryNameIndex({}).lookup(np.zeros(0, dtype="S1"), "x").tolist() == []


def _fake_results(rnd, nq, idx_ids, n_docs, recs):
    """(units, hits, n_kmers, _lib.Results) fabricated on the host (no GPU involved)."""
    import ctypes as C
    from phylign_b200 import _lib
    from phylign_b200.matcher import HIT_DT, UNIT_DT
    units, hits = [], []
    for i in idx_ids:
        for q in range(nq):
            if rnd.random() < 0.3 and recs[q][1]:
                k = rnd.randrange(1, 6)
                units.append((q, i, k + rnd.randrange(3), k, len(hits)))
                hits += [(rnd.randrange(n_docs), rnd.randrange(1, 99999)) for _ in range(k)]
    ua, ha = np.array(units, dtype=UNIT_DT), np.array(hits, dtype=HIT_DT)
    nk = np.zeros(nq, np.uint32)
    r = _lib.Results(nq, len(idx_ids), len(ua), C.cast(ua.ctypes.data, C.POINTER(_lib.Unit)), len(ha),
                     C.cast(ha.ctypes.data, C.POINTER(_lib.Hit)), C.cast(nk.ctypes.data, C.POINTER(C.c_uint32)), 0, 0)
    return ua, ha, nk, r


@pytest.mark.parametrize("nq,gz", [(50, 1), (20000, 1), (300, 0)])
def test_native_match_file_writer_equals_python_twin(tmp_path, nq, gz):
    """phy_write_match_blocks (threads + zlib, one gzip member per task) leaves files whose decompressed
    content is the twin formatter's text, block after block; nothing is visible before commit()."""
    import ctypes as C
    import random
    from phylign_b200.cobs_index import ClassicHeader
    from phylign_b200.cobs_text import _cat
    from phylign_b200.match_files import MatchFileSet
    from phylign_b200.matcher import MatchResult, ResidentIndex
    from text_twins import format_cobs_text
    rnd = random.Random(nq)
    n_docs = 41
    names = [f"{rnd.randrange(10**6):06d}_ACC{d:04d}" for d in range(n_docs)]
    ixs = {i: ResidentIndex(i, f"b__0{i}", ClassicHeader(31, 1, n_docs, 10, 1, names)) for i in (0, 3, 4)}
    paths = {i: str(tmp_path / f"b__0{i}____q{'.gz' if gz else '.txt'}") for i in ixs}
    fs = MatchFileSet(paths, ixs, gzip_level=gz, threads=5)
    want = {i: "" for i in ixs}
    for block in range(3):
        recs = [(f"q{block}_{q} c", "ACGT" * 10) for q in range(nq if block != 1 else 7)]
        ua, ha, nk, r = _fake_results(rnd, len(recs), list(ixs), n_docs, recs)
        res = MatchResult(ua, ha, nk, len(recs), 0, 0)
        hcat, hoffs = _cat([h for h, _ in recs])
        fs.write_block(hcat, hoffs, C.pointer(r))
        for i in ixs:
            want[i] += format_cobs_text(recs, res, ixs[i], strip_prefix=True)
    assert not any(os.path.exists(p) for p in paths.values())
    fs.commit()
    for i, p in paths.items():
        got = (gzip.open(p, "rt") if gz else open(p)).read()
        assert got == want[i]
        assert fs.file_bytes[i] == os.path.getsize(p)
    st = fs.stats_dict()
    assert st["header_lines"] == 3 * (2 * nq + 7) and st["text_bytes"] == sum(len(w) for w in want.values())
    assert not [f for f in os.listdir(tmp_path) if ".tmp." in f]
    # an aborted set leaves nothing behind; an empty committed gzip file is a valid gzip stream
    fs2 = MatchFileSet({0: str(tmp_path / "x.gz")}, ixs, gzip_level=1)
    fs2.abort()
    assert not os.path.exists(tmp_path / "x.gz") and not [f for f in os.listdir(tmp_path) if ".tmp." in f]
    fs3 = MatchFileSet({0: str(tmp_path / "e.gz")}, ixs, gzip_level=1)
    fs3.commit()
    assert gzip.open(tmp_path / "e.gz").read() == b""


def test_view_keeps_owner_alive():
    """Result arrays (and slices of them) hold the library block; it is released with the last view."""
    import ctypes as C
    import gc
    from phylign_b200.matcher import HIT_DT, _Owned, _view
    raw = (C.c_char * 64)()
    freed = []
    owner = _Owned(C.cast(raw, C.POINTER(C.c_char)), lambda p: freed.append(1))
    a = _view(C.cast(raw, C.POINTER(C.c_char)), 4, HIT_DT, owner)
    part = a[1:3]
    del owner, a
    gc.collect()
    assert not freed
    del part
    gc.collect()
    assert freed == [1]


def test_native_query_file_equals_python_reader(tmp_path):
    """phy_fasta_read (flat arrays) == read_cobs_records, incl. the `simple` verdict that lets match-db
    treat the record list as filter_queries.py's query dict."""
    from phylign_b200.fasta import QueryFile, read_cobs_records, read_fastx
    p = os.path.join(H.GOLDEN, "queries.fa")
    q = QueryFile(p)
    assert [(h, s.decode()) for h, s in q.records()] == read_cobs_records(p)
    assert q.simple and q.names() == [n for n, _ in read_fastx(p)]
    assert q.total_bases == sum(len(s) for _, s in read_cobs_records(p))
    assert q.block_ranges(10 ** 9) == [(0, q.n)]
    br = q.block_ranges(400)
    assert br[0][0] == 0 and br[-1][1] == q.n and all(a[1] == b[0] for a, b in zip(br, br[1:]))
    assert all(int(q.soffs[b] - q.soffs[a]) <= 400 or b == a + 1 for a, b in br)
    cases = {"semi": ">a desc here\nACGT\nAC\n\n;empty\n>b\nGG\n>c x\nTT\n",
             "crlf": ">a\r\nACGT\r\n>b\r\nGG\r\n", "noeol": ">a\nACGT\n>b\nGG", "fq": ">a\nAC\n+\nII\n",
             "empty_rec": ">a\n>b\nGG\n", "tabname": ">a\tx y\nACGT\n", "plain": ">a x\nAC\nGT\n>b\nTT\n"}
    for name, txt in cases.items():
        f = tmp_path / f"{name}.fa"
        f.write_bytes(txt.encode())
        q = QueryFile(str(f))
        assert [(h, s.decode()) for h, s in q.records()] == read_cobs_records(str(f)), name
        assert q.simple == (name == "plain"), name
    gz = tmp_path / "q.fa.gz"
    with gzip.open(gz, "wt") as f:
        f.write(cases["plain"])
    qz = QueryFile(str(gz))
    assert [(h, s.decode()) for h, s in qz.records()] == [("a x", "ACGT"), ("b", "TT")] and qz.names() == ["a", "b"]


def test_native_filter_fasta_writer_equals_twin(tmp_path):
    import ctypes as C
    import random
    from phylign_b200 import _lib
    from phylign_b200.cobs_text import write_filter_fasta_native
    from phylign_b200.fasta import QueryFile
    from phylign_b200.matcher import CAND_DT
    from text_twins import format_filter_fasta
    rnd = random.Random(5)
    nq, n_docs = 300, 23
    fa = tmp_path / "q.fa"
    fa.write_text("".join(f">q{q} comment {q}\n{'ACGT' * (1 + q % 7)}\n" for q in range(nq)))
    qf = QueryFile(str(fa))
    refs = {0: [f"SAM{d}" for d in range(n_docs)], 1: [], 2: [f"ERR{d}" for d in range(n_docs)]}
    offs = np.zeros(nq + 1, np.uint64)
    cands = []
    for q in range(nq):
        for _ in range(rnd.randrange(0, 6)):
            cands.append((rnd.randrange(1, 500), rnd.choice([0, 2]), rnd.randrange(n_docs), 0))
        offs[q + 1] = len(cands)
    ca = np.array(cands, dtype=CAND_DT)
    m = _lib.Merged(nq, C.cast(offs.ctypes.data, C.POINTER(C.c_uint64)), C.cast(ca.ctypes.data, C.POINTER(_lib.Cand)), 0)
    out = tmp_path / "04" / "q.fa"
    os.makedirs(out.parent)
    n = write_filter_fasta_native(str(out), C.pointer(m), qf, refs)
    want = format_filter_fasta([(f"q{q}", "ACGT" * (1 + q % 7)) for q in range(nq)], offs, ca, refs)
    assert out.read_text() == want and n == len(want)
    assert os.listdir(out.parent) == ["q.fa"]


def test_match_db_gpus_parent_stops_when_a_worker_dies(tmp_path):
    """`match-db --gpus N`: a worker that fails (here: no usable GPU, so every Matcher() raises) must take
    the whole job down with a non-zero exit instead of leaving the others blocked in a collective
    (the parent used to wait for every child in turn)."""
    import subprocess
    import sys
    import time
    batches = tmp_path / "batches.txt"
    batches.write_text("\n".join(H.GOLDEN_BATCHES) + "\n")
    t0 = time.time()
    r = subprocess.run([sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", H.GOLDEN, "--batches",
                        str(batches), "-q", os.path.join(H.GOLDEN, "queries.fa"), "--match-dir", str(tmp_path / "m"),
                        "--filter-out", str(tmp_path / "f.fa"), "--gpus", "2"], capture_output=True, text=True,
                       cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES=""), timeout=120)
    assert r.returncode != 0 and "worker" in r.stderr and "failed" in r.stderr, r.stderr[-1000:]
    assert time.time() - t0 < 60
    assert not os.path.exists(tmp_path / "f.fa") and not [f for f in os.listdir(tmp_path) if ".part." in f]


def test_postprocess_keep_zero_and_short_blocks():
    """`postprocess -n N` on blocks shorter than N, N = 1 ties and the N = 0 quirk."""
    import io
    from oracle import filters
    from phylign_b200.cli import _postprocess_stream
    txt = "*q1\t5\nzz9_A\t6\nab1_B\t6\nqq2_C\t5\nqq3_D\t5\nqq4_E\t4\n*q2\t0\n*q3\t1\nxx_F\t8\n"
    for keep in (0, 1, 2, 3, 4, 5, 9):
        out = io.StringIO()
        _postprocess_stream(io.StringIO(txt), out, keep)
        if keep >= 1:
            assert out.getvalue() == filters.postprocess_text(txt, keep), keep
    out = io.StringIO()
    _postprocess_stream(io.StringIO(txt), out, 1)
    assert out.getvalue() == "*q1\t5\n_A\t6\n_B\t6\n*q2\t0\n*q3\t1\n_F\t8\n"


def test_benchmark_log_has_the_reference_columns(tmp_path):
    """--benchmark-dir files start with the header scripts/benchmark.py writes (its :33-46), so the
    reference's log readers find the 8 columns they know before the match-stage columns."""
    from phylign_b200.cli import _write_benchmark_log
    p = tmp_path / "logs" / "benchmarks" / "run_cobs" / "b__01____q.txt"
    _write_benchmark_log(str(p), "phylign_b200 match-db (round 0)", 1.25, [("batch", "b__01"), ("gpu_ms", "3.5")])
    lines = p.read_text().splitlines()
    assert lines[0].startswith("# Benchmarking command: ")
    head, vals = lines[1].split("\t"), lines[2].split("\t")
    assert head[:8] == ["real(s)", "sys(s)", "user(s)", "percent_CPU", "max_RAM(kb)", "FS_inputs", "FS_outputs",
                        "elapsed_time_alt(s)"]
    assert len(head) == len(vals) == 10 and float(vals[0]) == 1.25 and vals[8] == "b__01"


def test_join_parts_orders_by_block_then_rank(tmp_path):
    from phylign_b200.cli import _join_parts
    out = tmp_path / "q.fa"
    for name, txt in (("000001.0000", "C"), ("000000.0001", "B"), ("000000.0000", "A"), ("000001.0001", "D")):
        (tmp_path / f"q.fa.part.{name}").write_text(txt)
    _join_parts(str(out), [str(p) for p in tmp_path.iterdir()])
    assert out.read_text() == "ABCD" and os.listdir(tmp_path) == ["q.fa"]
    (tmp_path / "q.fa.part.000000.0000").write_text("only")
    _join_parts(str(out), [str(tmp_path / "q.fa.part.000000.0000")])
    assert out.read_text() == "only" and os.listdir(tmp_path) == ["q.fa"]


def test_query_block_ranges_cover_every_record_once():
    """QueryFile.block_ranges: consecutive, non-empty, within the base budget unless a single record exceeds it."""
    import random
    import tempfile
    from phylign_b200.fasta import QueryFile
    rnd = random.Random(4)
    for _ in range(20):
        lens = [rnd.choice([1, 31, 150, 1000, 5000]) for _ in range(rnd.randrange(1, 60))]
        with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
            f.write("".join(f">r{i}\n{'A' * n}\n" for i, n in enumerate(lens)))
        q = QueryFile(f.name)
        os.unlink(f.name)
        budget = rnd.choice([1, 100, 1000, 6000, 10 ** 9])
        br = q.block_ranges(budget)
        assert br[0][0] == 0 and br[-1][1] == len(lens)
        for (a, b), nxt in zip(br, br[1:] + [(len(lens), None)]):
            assert b > a and b == nxt[0]
            assert sum(lens[a:b]) <= budget or b == a + 1


def test_native_reader_equals_python_reader_on_random_text(tmp_path):
    """phy_fasta_read vs read_cobs_records on 300 random line soups ('>' / ';' headers, blank and CRLF lines,
    text before the first header, missing final newline, '@' / '+' lines)."""
    import random
    from phylign_b200.fasta import QueryFile, read_cobs_records
    rnd = random.Random(12)
    tokens = [">q", ">q x y", ";c", "ACGT", "acgtn", "", "\r", "@r", "+", "GG\r", ">", "TTTTTTTTTT" * 8, " ", ">q\tz"]
    for case in range(300):
        lines = [rnd.choice(tokens) + (str(rnd.randrange(9)) if rnd.random() < 0.3 else "") for _ in range(rnd.randrange(0, 25))]
        txt = "\n".join(lines) + ("\n" if rnd.random() < 0.8 else "")
        p = tmp_path / "r.fa"
        p.write_bytes(txt.encode())
        want = read_cobs_records(str(p))
        q = QueryFile(str(p))
        got = [(h, s.decode()) for h, s in q.records()]
        assert got == want, (case, txt)
        assert q.n == len(want) and q.total_bases == sum(len(s) for _, s in want)
        assert q.names() == [h.split(" ")[0] for h, _ in want]


_FAKE_WORKERS = r"""
import os, subprocess, sys
sys.path.insert(0, {root!r})
import phylign_b200.cli as cli
real = subprocess.Popen
worker = ("import os, sys, time; sys.path.insert(0, %r); import phylign_b200.cli as c; c._die_with_parent(); "
          "open(os.environ['PIDF'] + '.' + os.environ['PHYLIGN_RANK'], 'w').write(str(os.getpid())); time.sleep(120)" % {root!r})
subprocess.Popen = lambda cmd, env=None: real([sys.executable, "-c", worker], env=env)
sys.argv = ["cli", "match-db", "--cobs-dir", "x", "--batches", "x", "-q", "x", "--match-dir", "x", "--gpus", "2"]
cli.main()
"""


def _alive(pid):
    try:
        with open(f"/proc/{pid}/stat") as f:
            return f.read().split(")")[-1].split()[0] != "Z"
    except OSError:
        return False


def _run_fake_job(tmp_path, sig):
    """Parent of `match-db --gpus 2` with stand-in workers (sleepers that call _die_with_parent), hit by `sig`."""
    import signal
    import subprocess
    import sys
    import time
    pidf = str(tmp_path / "pid")
    p = subprocess.Popen([sys.executable, "-c", _FAKE_WORKERS.format(root=ROOT)], env=dict(os.environ, PIDF=pidf))
    deadline = time.time() + 60
    while not (os.path.exists(pidf + ".0") and os.path.exists(pidf + ".1")
               and open(pidf + ".0").read() and open(pidf + ".1").read()):
        assert time.time() < deadline and p.poll() is None, "workers did not start"
        time.sleep(0.05)
    pids = [int(open(f"{pidf}.{r}").read()) for r in (0, 1)]
    assert all(_alive(w) for w in pids)
    p.send_signal(sig)
    rc = p.wait(timeout=30)
    deadline = time.time() + 10
    while any(_alive(w) for w in pids) and time.time() < deadline:
        time.sleep(0.05)
    return rc, [w for w in pids if _alive(w)]


def test_match_db_gpus_cancelled_parent_takes_its_workers_along(tmp_path):
    """SIGTERM to the supervising parent (a cancelled snakemake job) must not leave workers behind holding
    their GPUs inside a collective."""
    import signal
    rc, left = _run_fake_job(tmp_path, signal.SIGTERM)
    assert rc == 128 + signal.SIGTERM and not left


def test_match_db_gpus_workers_die_with_a_killed_parent(tmp_path):
    """SIGKILL gives the parent no chance to clean up: the workers' PR_SET_PDEATHSIG ends them."""
    import signal
    rc, left = _run_fake_job(tmp_path, signal.SIGKILL)
    assert rc == -signal.SIGKILL and not left
