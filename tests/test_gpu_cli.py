"""The command-line drop-ins end to end on the GPU, against the golden vectors."""
import gzip
import os
import subprocess
import sys

import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = dict(os.environ, PYTHONPATH=ROOT)


def _run(args, **kw):
    return subprocess.run(args, capture_output=True, text=True, env=ENV, cwd=ROOT, **kw)


@pytest.mark.parametrize("batch", H.GOLDEN_BATCHES)
def test_run_cobs_streaming_dropin(batch):
    xz = os.path.join(H.GOLDEN, f"{batch}.cobs_classic.xz")
    size = len(H.golden_index_bytes(batch))
    r = _run([os.path.join(ROOT, "scripts", "run_cobs_streaming.sh"), "0.7", "4", xz, str(size),
              os.path.join(H.GOLDEN, "queries.fa")])
    assert r.returncode == 0, r.stderr
    assert r.stdout == H.golden_cobs_text(batch)
    # wrong --index-sizes must fail loudly
    r = _run([os.path.join(ROOT, "scripts", "run_cobs_streaming.sh"), "0.7", "4", xz, str(size + 1),
              os.path.join(H.GOLDEN, "queries.fa")])
    assert r.returncode != 0 and "index-sizes" in r.stderr


def test_cobs_query_reads_index_from_pipe_and_feeds_reference_postprocess(tmp_path):
    """`cobs query -i <(xzcat ...)` exactly as run_cobs_streaming.sh:24-29 spells it."""
    batch = "bbb__02"
    xz = os.path.join(H.GOLDEN, f"{batch}.cobs_classic.xz")
    cmd = (f"{ROOT}/scripts/cobs query --load-complete -t 0.7 -T 8 "
           f"-i <(xzcat --no-sparse --ignore-check {xz}) --index-sizes {len(H.golden_index_bytes(batch))} "
           f"-f {H.GOLDEN}/queries.fa | {sys.executable} -m phylign_b200.cli postprocess -n 3")
    r = _run(["bash", "-o", "pipefail", "-c", cmd])
    assert r.returncode == 0, r.stderr
    assert r.stdout == H.golden_match_text(batch, 3)
    r2 = _run([f"{ROOT}/scripts/cobs", "query", "-t", "0.7", "-i", xz, "-f", f"{H.GOLDEN}/queries.fa",
               "--top-n", "3"])
    assert r2.returncode == 0 and r2.stdout == H.golden_match_text(batch, 3)


@pytest.mark.parametrize("keep", [1, 3, 100])
def test_filter_dropin_on_reference_match_files(keep):
    files = [os.path.join(H.GOLDEN, f"n{keep}", f"{b}____queries.gz") for b in H.GOLDEN_BATCHES]
    r = _run([sys.executable, os.path.join(ROOT, "scripts", "filter_queries_gpu.py"), "-n", str(keep),
              "-q", os.path.join(H.GOLDEN, "queries.fa")] + files[::-1])
    assert r.returncode == 0, r.stderr
    assert r.stdout == H.golden_filter_fa(keep)


def test_match_db_writes_all_rule_outputs(tmp_path):
    batches = tmp_path / "batches.txt"
    batches.write_text("\n".join(reversed(H.GOLDEN_BATCHES)) + "\n\n")
    mdir, out = tmp_path / "03_match", tmp_path / "04_filter" / "queries.fa"
    r = _run([sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", H.GOLDEN, "--batches",
              str(batches), "-q", os.path.join(H.GOLDEN, "queries.fa"), "--match-dir", str(mdir),
              "--filter-out", str(out), "-t", "0.7", "-n", "3"])
    assert r.returncode == 0, r.stderr
    for b in H.GOLDEN_BATCHES:
        got = gzip.open(mdir / f"{b}____queries.gz", "rt").read()
        assert got == H.golden_match_text(b, 3)
    assert out.read_text() == H.golden_filter_fa(3)
    assert not [f for f in os.listdir(mdir) if ".tmp." in f]
    # 04 -> 05 hand-off tables: one per batch, every candidate of the FASTA appears exactly once
    bdir = tmp_path / "buckets"
    r = _run([sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", H.GOLDEN, "--batches",
              str(batches), "-q", os.path.join(H.GOLDEN, "queries.fa"), "--match-dir", str(tmp_path / "m2"),
              "--filter-out", str(tmp_path / "f2.fa"), "--bucket-dir", str(bdir), "-t", "0.7", "-n", "3"])
    assert r.returncode == 0, r.stderr
    pairs = set()
    for b in H.GOLDEN_BATCHES:
        for line in open(bdir / f"{b}____queries.candidates.tsv"):
            ref, qs = line.rstrip("\n").split("\t")
            pairs |= {(q, ref) for q in qs.split(",")}
    want = set()
    for line in H.golden_filter_fa(3).splitlines():
        if line.startswith(">"):
            name, _, com = line[1:].partition(" ")
            want |= {(name, a) for a in com.split(",") if a}
    assert pairs == want
    # a missing batch fails, and leaves no partial file behind
    batches.write_text("nosuch__01\n")
    r = _run([sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", H.GOLDEN, "--batches",
              str(batches), "-q", os.path.join(H.GOLDEN, "queries.fa"), "--match-dir", str(tmp_path / "x")])
    assert r.returncode != 0 and not os.listdir(tmp_path / "x")


def test_match_db_shards_cover_all_batches(tmp_path):
    """Two `--shard i/2` processes (one per GPU in production) write disjoint match files whose
    union is complete; the reference's own merge step then reproduces 04_filter."""
    batches = tmp_path / "batches.txt"
    batches.write_text("\n".join(H.GOLDEN_BATCHES) + "\n")
    mdir = tmp_path / "03_match"
    for i in range(2):
        r = _run([sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", H.GOLDEN, "--batches",
                  str(batches), "-q", os.path.join(H.GOLDEN, "queries.fa"), "--match-dir", str(mdir),
                  "-t", "0.7", "-n", "100", "--shard", f"{i}/2"])
        assert r.returncode == 0, r.stderr
        assert 1 <= len(os.listdir(mdir)) <= 3
    assert sorted(os.listdir(mdir)) == sorted(f"{b}____queries.gz" for b in H.GOLDEN_BATCHES)
    files = [str(mdir / f"{b}____queries.gz") for b in H.GOLDEN_BATCHES]
    r = _run([sys.executable, os.path.join(ROOT, "scripts", "filter_queries_gpu.py"), "-n", "100",
              "-q", os.path.join(H.GOLDEN, "queries.fa")] + files)
    assert r.returncode == 0 and r.stdout == H.golden_filter_fa(100)
    # resume: nothing to do, outputs untouched
    before = {f: os.path.getmtime(mdir / f) for f in os.listdir(mdir)}
    r = _run([sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", H.GOLDEN, "--batches",
              str(batches), "-q", os.path.join(H.GOLDEN, "queries.fa"), "--match-dir", str(mdir),
              "-t", "0.7", "-n", "100", "--resume"])
    assert r.returncode == 0 and before == {f: os.path.getmtime(mdir / f) for f in os.listdir(mdir)}


def test_match_db_streams_overflow_rounds(tmp_path):
    """Indexes that do not fit together are streamed through HBM in rounds (load, match, write,
    evict); the outputs do not depend on the round structure."""
    batches = tmp_path / "batches.txt"
    batches.write_text("\n".join(H.GOLDEN_BATCHES) + "\n")
    mdir, out = tmp_path / "03_match", tmp_path / "q.fa"
    r = _run([sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", H.GOLDEN, "--batches",
              str(batches), "-q", os.path.join(H.GOLDEN, "queries.fa"), "--match-dir", str(mdir),
              "--filter-out", str(out), "-t", "0.7", "-n", "1", "--round-bytes", "500000"])
    assert r.returncode == 0, r.stderr
    for b in H.GOLDEN_BATCHES:
        assert gzip.open(mdir / f"{b}____queries.gz", "rt").read() == H.golden_match_text(b, 1)
    assert out.read_text() == H.golden_filter_fa(1)
    # a batch larger than the round budget is an error, not a silent skip
    r = _run([sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", H.GOLDEN, "--batches",
              str(batches), "-q", os.path.join(H.GOLDEN, "queries.fa"), "--match-dir", str(tmp_path / "y"),
              "--round-bytes", "1000"])
    assert r.returncode != 0 and "budget" in r.stderr


def test_resident_server_answers_cobs_queries(tmp_path):
    """`serve` keeps indexes in HBM; `cobs query --server` prints the same bytes as a cold run."""
    import time
    from phylign_b200.server import request
    sock = str(tmp_path / "phy.sock")
    srv = subprocess.Popen([sys.executable, "-m", "phylign_b200.cli", "serve", "--socket", sock,
                            "--preload", os.path.join(H.GOLDEN, "aaa__01.cobs_classic.xz")],
                           env=ENV, cwd=ROOT, stderr=subprocess.PIPE, text=True)
    try:
        for _ in range(600):
            if os.path.exists(sock):
                break
            assert srv.poll() is None, srv.stderr.read()
            time.sleep(0.1)
        for rep in range(2):
            for b in H.GOLDEN_BATCHES:
                r = _run([f"{ROOT}/scripts/cobs", "query", "-t", "0.7", "-T", "4", "-i",
                          os.path.join(H.GOLDEN, f"{b}.cobs_classic.xz"), "-f", f"{H.GOLDEN}/queries.fa",
                          "--server", sock])
                assert r.returncode == 0, r.stderr
                assert r.stdout == H.golden_cobs_text(b), (rep, b)
        head, _ = request(sock, {"cmd": "status"})
        assert head["ok"] and head["loads"] == 3 and head["hits"] == 4 and head["queries"] == 6
        r = _run([f"{ROOT}/scripts/cobs", "query", "-t", "0.7", "-i", "/nonexistent.cobs_classic",
                  "-f", f"{H.GOLDEN}/queries.fa", "--server", sock])
        assert r.returncode != 0 and "error" in r.stderr
    finally:
        try:
            request(sock, {"cmd": "shutdown"}, timeout=30)
        except Exception:
            pass
        try:
            srv.wait(timeout=30)
        except Exception:
            srv.kill()


def test_query_blocks_do_not_change_the_output(tmp_path):
    """Queries processed in several HBM-sized blocks: same bytes as in one block."""
    xz = os.path.join(H.GOLDEN, "aaa__01.cobs_classic.xz")
    r = _run([f"{ROOT}/scripts/cobs", "query", "-t", "0.7", "-i", xz, "-f", f"{H.GOLDEN}/queries.fa",
              "--query-block-bases", "700"])
    assert r.returncode == 0 and r.stdout == H.golden_cobs_text("aaa__01"), r.stderr
    batches = tmp_path / "batches.txt"
    batches.write_text("\n".join(H.GOLDEN_BATCHES) + "\n")
    mdir, out = tmp_path / "03_match", tmp_path / "q.fa"
    r = _run([sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", H.GOLDEN, "--batches",
              str(batches), "-q", os.path.join(H.GOLDEN, "queries.fa"), "--match-dir", str(mdir),
              "--filter-out", str(out), "-t", "0.7", "-n", "3", "--query-block-bases", "1000",
              "--round-bytes", "500000"])
    assert r.returncode == 0, r.stderr
    for b in H.GOLDEN_BATCHES:
        assert gzip.open(mdir / f"{b}____queries.gz", "rt").read() == H.golden_match_text(b, 3)
    assert out.read_text() == H.golden_filter_fa(3)


def test_query_prep_on_device_equals_snakefile_rule(tmp_path):
    """phy_fix_bases / `fix-query --device` / `match-db --sanitize-queries` == rule fix_query
    (Snakefile:326-332: seqtk seq -A -U -C | awk gsub(/[^ACGT]/,"A")) as restated on the host."""
    import numpy as np
    from phylign_b200.fasta import fix_query_seq
    from phylign_b200.matcher import Matcher
    raw = bytes(range(256)) * 9 + b"acgtnNRYKMswbdhv-*ACGT" * 41 + b"x"           # every byte value, odd length
    with Matcher(0) as m:
        arr = np.frombuffer(bytearray(raw), dtype=np.uint8)
        m.fix_bases(arr)
        assert arr.tobytes().decode("latin1") == fix_query_seq(raw.decode("latin1").encode("latin1"))
        assert set(arr.tobytes()) <= set(b"ACGT")
    fq = tmp_path / "r.fq"
    fq.write_text("@r1 some comment\nacgtnACGT\n+\nIIIIIIIII\n@r2\nGGNN\n+\nIIII\n")
    fa = tmp_path / "g.fa"
    fa.write_text(">g1 desc\nACGT\nryk\n>g2\nTTTT\n")
    want = _run([sys.executable, "-m", "phylign_b200.cli", "fix-query", str(fq), str(fa)])
    got = _run([sys.executable, "-m", "phylign_b200.cli", "fix-query", "--device", "0", str(fq), str(fa)])
    assert want.returncode == 0 and got.returncode == 0, got.stderr
    assert got.stdout == want.stdout == ">r1\nACGTAACGT\n>r2\nGGAA\n>g1\nACGTAAA\n>g2\nTTTT\n"
    # raw (lower-case, N-containing) queries through match-db --sanitize-queries == sanitised queries through match-db
    recs = H.read_fasta(os.path.join(H.GOLDEN, "queries.fa"))
    dirty = tmp_path / "dirty.fa"
    dirty.write_text("".join(f">{h}\n{s.lower() if i % 2 else s}\n" for i, (h, s) in enumerate(recs)))
    batches = tmp_path / "batches.txt"
    batches.write_text("\n".join(H.GOLDEN_BATCHES) + "\n")
    r = _run([sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", H.GOLDEN, "--batches", str(batches),
              "-q", str(dirty), "--qfile", "queries", "--match-dir", str(tmp_path / "m"), "--filter-out",
              str(tmp_path / "f.fa"), "-t", "0.7", "-n", "3", "--sanitize-queries"])
    assert r.returncode == 0, r.stderr
    for b in H.GOLDEN_BATCHES:
        assert gzip.open(tmp_path / "m" / f"{b}____queries.gz", "rt").read() == H.golden_match_text(b, 3)
    assert (tmp_path / "f.fa").read_text() == H.golden_filter_fa(3)
    # without the flag the dirty file is rejected (cobs aborts on such letters too)
    r = _run([sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", H.GOLDEN, "--batches", str(batches),
              "-q", str(dirty), "--qfile", "queries", "--match-dir", str(tmp_path / "m2"), "-t", "0.7", "-n", "3"])
    assert r.returncode != 0 and "ACGT" in r.stderr


def test_pipeline_fuzz_a_few_cases():
    """tests/fuzz_pipeline.py (random databases + query files through match-db with random blocking /
    rounds / resume, against cobs_oracle | postprocess_cobs.py | gzip -> filter_queries.py): 6 cases here,
    40 on 2 GPUs in profiles/r02_fuzz_pipeline_vs_reference_scripts.txt."""
    r = _run([sys.executable, os.path.join(ROOT, "tests", "fuzz_pipeline.py"), "6", "7"], timeout=900)
    assert r.returncode == 0 and "6 random databases" in r.stdout, (r.stdout + r.stderr)[-3000:]
