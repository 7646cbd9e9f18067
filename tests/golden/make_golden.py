#!/usr/bin/env python3
"""Generate tests/golden/* by running the UNMODIFIED reference scripts.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py [--with-cobs]

--with-cobs: when an executable `cobs` (COBS 0.2.1, envs/cobs.yaml:5) is on PATH or in
$PHYLIGN_REAL_COBS, the three indexes are ALSO built with `cobs classic-construct` and must equal the
oracle's byte for byte, and the <batch>.cobs.txt.gz texts are taken from the real
`cobs query --load-complete -t 0.7 -T 2 -i ... -f ...` (run_cobs_streaming.sh:24-29; equal-score lines
put in document order, which cobs leaves undefined).  PROVENANCE.json records which producer wrote the
vectors, so "parity unpinned" can be dropped from the docs the day this has run.

What is produced (all small, committed):
  xxh64_kat.json            XXH64 known answers from python-xxhash AND libxxhash
                            (two independent implementations must agree)
  queries.fa                the reference's 40 test reads after the fix_query
                            transform (/root/reference/Snakefile:326-332) plus
                            synthetic reads and edge-case records
  <batch>.cobs_classic.xz   three small synthetic classic indexes (oracle
                            classic-construct restatement; synthetic spec v1)
  <batch>.cobs.txt.gz       oracle `cobs query -t 0.7` text for each batch
  n<N>/<batch>____queries.gz   reference postprocess_cobs.py -n N | gzip output
  n<N>/queries.fa              reference filter_queries.py -n N output
The filter outputs are TRUE reference behaviour (the scripts are executed as
they are); the cobs text is the oracle's (parity unpinned vs. real cobs).
"""
import ctypes
import gzip
import json
import lzma
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

import oracle  # noqa: E402

BATCHES = [
    # name, seed, n_docs, genome_len
    ("aaa__01", 11, 200, 5000),
    ("bbb__01", 22, 37, 4000),
    ("bbb__02", 11, 130, 5000),  # same seed as aaa__01 -> same genomes -> cross-batch ties
]
THRESHOLD = 0.7
KEEPS = [1, 3, 100]


def spec_of(seed, n_docs, genome_len):
    return oracle.SynthSpec(seed=seed, n_docs=n_docs, genome_len=genome_len, clade_size=16,
                            clade_sub_q16=328, doc_sub_q16=328)


def doc_names(batch_i, n_docs):
    rnd = random.Random(1000 + batch_i)
    prefixes = sorted(rnd.sample(range(10 ** 6), n_docs))
    accs = list(range(n_docs))
    rnd.shuffle(accs)  # accession order differs from doc (prefix) order
    return [f"{p:06d}_SAMS{batch_i:02d}{a:05d}" for p, a in zip(prefixes, accs)]


def read_fastx(path):
    recs = []
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f]
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith(">"):
            name = ln[1:].split()[0]
            i += 1
            seq = []
            while i < len(lines) and not lines[i].startswith(">"):
                seq.append(lines[i])
                i += 1
            recs.append((name, "".join(seq)))
        elif ln.startswith("@"):
            recs.append((ln[1:].split()[0], lines[i + 1]))
            i += 4
        else:
            i += 1
    return recs


def fix_query(seq):
    """seqtk seq -U + awk non-ACGT -> A (/root/reference/Snakefile:326-332)."""
    return "".join(c if c in "ACGT" else "A" for c in seq.upper())


def xxh64_kats():
    import xxhash
    libxx = ctypes.CDLL("libxxhash.so.0")
    libxx.XXH64.restype = ctypes.c_uint64
    libxx.XXH64.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64]
    rnd = random.Random(7)
    vecs = []
    inputs = [b"", b"A", b"A" * 31, b"ACGTACGTACGTACGTACGTACGTACGTACG",
              b"ATTCTGATGAGGGCTTATTACATGAACCAAT", b"A" * 32, b"ACGT" * 25]
    for ln in (3, 4, 7, 8, 15, 16, 20, 30, 31, 32, 33, 63, 64, 95, 127):
        inputs.append(bytes(rnd.choice(b"ACGT") for _ in range(ln)))
    for _ in range(40):
        inputs.append(bytes(rnd.choice(b"ACGT") for _ in range(31)))
    for data in inputs:
        for seed in (0, 1, 2, 0xDEADBEEF):
            a = xxhash.xxh64(data, seed=seed).intdigest()
            b = libxx.XXH64(data, len(data), seed)
            assert a == b, (data, seed)
            vecs.append({"data": data.decode(), "seed": seed, "hash": f"{a:016x}"})
    return vecs


def real_cobs():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_real_cobs import real_cobs as find
    return find()


def cobs_construct(cobs, names, docs):
    """`cobs classic-construct` over one FASTA per document (file stem = document name)."""
    with tempfile.TemporaryDirectory() as td:
        ddir = os.path.join(td, "docs")
        os.makedirs(ddir)
        for n, s in zip(names, docs):
            with open(os.path.join(ddir, n + ".fa"), "w") as f:
                f.write(f">{n}\n{s.decode()}\n")
        out = os.path.join(td, "built.cobs_classic")
        base = [cobs, "classic-construct", "-k", "31", "--num-hashes", "1", "--false-positive-rate", "0.3"]
        for variant in (["--clobber"], []):
            r = subprocess.run(base + variant + [ddir, out], capture_output=True, text=True)
            if r.returncode == 0 and os.path.exists(out):
                return open(out, "rb").read()
        raise SystemExit("cobs classic-construct failed: " + r.stderr[-500:])


def canonical_tie_order(text, names):
    """cobs text with the equal-score lines of every block in document order."""
    from oracle import filters
    pos = {n: i for i, n in enumerate(names)}
    out = []
    for head, n, hits in filters.parse_cobs_text(text):
        out.append(f"*{head}\t{n}\n")
        out.extend(f"{nm}\t{sc}\n" for nm, sc in sorted(hits, key=lambda x: (-x[1], pos[x[0]])))
    return "".join(out)


def main():
    with_cobs = "--with-cobs" in sys.argv
    cobs = real_cobs() if with_cobs else None
    if with_cobs and not cobs:
        raise SystemExit("--with-cobs: no `cobs` executable on PATH and $PHYLIGN_REAL_COBS unset")
    oracle.build()
    with open(os.path.join(HERE, "xxh64_kat.json"), "w") as f:
        json.dump(xxh64_kats(), f, indent=0)

    specs = [spec_of(s, n, g) for _, s, n, g in BATCHES]
    # ---- queries --------------------------------------------------------
    queries = []
    for fn in ("reads_1.fastq", "reads_2.fq", "reads_3.fasta", "reads_4.fa"):
        for name, seq in read_fastx(os.path.join(REF, "data", fn)):
            queries.append((name, fix_query(seq)))
    for r in range(40):
        queries.append((f"syn{r:02d}", oracle.synth_read(specs, 99, r, 150, random_q8=26,
                                                         err_q16=655).decode()))
    g0 = oracle.synth_genome(specs[0], 5).decode()
    queries.append(("exact_doc5 with a comment", g0[100:400]))       # comment must be dropped downstream
    queries.append(("len31", g0[1000:1031]))
    queries.append(("len40", g0[2000:2040]))
    queries.append(("short30", g0[0:30]))                              # L<k -> 0 terms
    queries.append(("polyA", "A" * 120))
    queries.append(("long2k", g0[500:2600]))                           # K=2070 > 1023
    with open(os.path.join(HERE, "queries.fa"), "w") as f:
        for name, seq in queries:
            f.write(f">{name}\n{seq}\n")

    # ---- indexes + oracle cobs text ----------------------------------------
    for bi, ((batch, seed, n_docs, glen), spec) in enumerate(zip(BATCHES, specs)):
        docs = [oracle.synth_genome(spec, d) for d in range(n_docs)]
        idx = oracle.OracleIndex.construct(docs, doc_names(bi, n_docs))
        with tempfile.TemporaryDirectory() as td:
            p = os.path.join(td, "i.cobs_classic")
            idx.write(p)
            raw = open(p, "rb").read()
            if cobs:
                built = cobs_construct(cobs, doc_names(bi, n_docs), docs)
                if built != raw:
                    raise SystemExit(f"{batch}: `cobs classic-construct` output differs from the oracle's index "
                                     f"({len(built)} vs {len(raw)} bytes): fix the oracle (SURVEY Appendix A.1/A.10)")
            with lzma.open(os.path.join(HERE, f"{batch}.cobs_classic.xz"), "wb", preset=6) as f:
                f.write(raw)
            # the CLI and the python binding must agree
            txt_cli = subprocess.check_output(
                [oracle.CLI_PATH, "query", "--load-complete", "-t", str(THRESHOLD), "-T", "2",
                 "-i", p, "-f", os.path.join(HERE, "queries.fa")]).decode()
        txt = idx.query_text([(n, s.encode()) for n, s in queries], THRESHOLD)
        assert txt == txt_cli
        if cobs:        # the vector comes from the real binary; the oracle must agree on everything it defines
            with tempfile.TemporaryDirectory() as td:
                p = os.path.join(td, "i.cobs_classic")
                open(p, "wb").write(raw)
                qok = os.path.join(td, "q.fa")       # (records shorter than k are an open item: test_real_cobs.py)
                with open(qok, "w") as f:
                    f.write("".join(f">{n}\n{s}\n" for n, s in queries if len(s) >= 31))
                real = subprocess.run([cobs, "query", "--load-complete", "-t", str(THRESHOLD), "-T", "2", "-i", p, "-f", qok],
                                      check=True, stdout=subprocess.PIPE).stdout.decode()
            names = doc_names(bi, n_docs)
            want = canonical_tie_order(idx.query_text([(n, s.encode()) for n, s in queries if len(s) >= 31], THRESHOLD), names)
            if canonical_tie_order(real, names) != want:
                raise SystemExit(f"{batch}: `cobs query` output differs from the oracle's: run tests/test_real_cobs.py "
                                 f"for the item-by-item diagnosis")
        with gzip.GzipFile(os.path.join(HERE, f"{batch}.cobs.txt.gz"), "wb", mtime=0) as f:
            f.write(txt.encode())

    # ---- reference filters, executed unmodified ----------------------------------
    shim = tempfile.mkdtemp()
    with open(os.path.join(shim, "xopen.py"), "w") as f:   # test-only stand-in for xopen 0.7.3
        f.write("import gzip\n"
                "def xopen(fn, mode='r'):\n"
                "    fn = str(fn)\n"
                "    return gzip.open(fn, mode + 't') if fn.endswith('.gz') else open(fn, mode)\n")
    env = dict(os.environ, PYTHONPATH=shim)
    for keep in KEEPS:
        od = os.path.join(HERE, f"n{keep}")
        os.makedirs(od, exist_ok=True)
        match_files = []
        for batch, *_ in BATCHES:
            txt = gzip.open(os.path.join(HERE, f"{batch}.cobs.txt.gz")).read()
            post = subprocess.run([sys.executable, os.path.join(REF, "scripts/postprocess_cobs.py"),
                                   "-n", str(keep)], input=txt, stdout=subprocess.PIPE, check=True).stdout
            mf = os.path.join(od, f"{batch}____queries.gz")
            with gzip.GzipFile(mf, "wb", mtime=0) as f:
                f.write(post)
            match_files.append(mf)
        fa = subprocess.run([sys.executable, os.path.join(REF, "scripts/filter_queries.py"), "-n", str(keep),
                             "-q", os.path.join(HERE, "queries.fa")] + match_files,
                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True, env=env).stdout
        with open(os.path.join(od, "queries.fa"), "wb") as f:
            f.write(fa)
    with open(os.path.join(HERE, "PROVENANCE.json"), "w") as f:
        json.dump({"cobs_text_and_indexes": ("real cobs binary: " + cobs + " (indexes byte-identical to the oracle's, query text "
                                             "identical up to the order of equal-score lines)") if cobs else
                   "oracle/cobs_oracle.c (parity with the real cobs binary UNPINNED: cobs not installable offline)",
                   "postprocess_and_filter_outputs": "unmodified /root/reference/scripts/postprocess_cobs.py and filter_queries.py, executed"},
                  f, indent=1)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
