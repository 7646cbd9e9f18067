#!/usr/bin/env python3
"""Randomised end-to-end parity of the drop-in against the reference PIPELINE (run on the GPU box).

Per case: a random small database (2-6 batches of random document counts / hash counts, names with
random sorting prefixes) and a random query FASTA (multi-line records, comments in headers, ';'
headers, records without sequence, sometimes duplicate names) go through
  (a) the CPU pipeline the Snakefile runs (Snakefile:416-428, 513-520):
        cobs_oracle query -t T | postprocess_cobs.py -n N | gzip  per batch, then filter_queries.py
      (the UNMODIFIED reference scripts from baseline/_ref/scripts when installed, else oracle/filters.py)
  (b) `python -m phylign_b200.cli match-db` with random --query-block-bases / --round-bytes /
      --hbm-budget / --gpus (when more than one GPU is visible) / --resume after a partial run.
Compared: every 03_match file (decompressed; equal-score lines as sets -- cobs leaves their order
undefined) and the 04_filter FASTA byte for byte.

    python tests/fuzz_pipeline.py [cases] [seed]
"""
import gzip
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle                                  # noqa: E402
from oracle import filters                     # noqa: E402

REF = os.path.join(ROOT, "baseline", "_ref", "scripts")
HAVE_REF = os.path.exists(os.path.join(REF, "filter_queries.py"))
ENV = dict(os.environ, PYTHONPATH=ROOT)


def n_gpus():
    import ctypes as C
    from phylign_b200 import _lib
    n = C.c_int()
    _lib.load().phy_device_count(C.byref(n))
    return n.value


def canon_match_text(text):
    """Blocks with equal-score lines sorted by name (their order is not defined by cobs)."""
    return [(h, n, sorted(hits, key=lambda x: (-x[1], x[0]))) for h, n, hits in filters.parse_cobs_text(text)]


def cpu_pipeline(td, cobs_dir, batches, qfa, thr, keep):
    """The reference pipeline on the CPU; returns ({batch: match text}, 04_filter text)."""
    mdir = os.path.join(td, "ref_03")
    os.makedirs(mdir, exist_ok=True)
    files, texts = [], {}
    for b in batches:
        raw = subprocess.run([oracle.CLI_PATH, "query", "--load-complete", "-t", str(thr), "-T", "2", "-i",
                              os.path.join(cobs_dir, f"{b}.cobs_classic"), "-f", qfa], check=True,
                             stdout=subprocess.PIPE).stdout
        if HAVE_REF:
            post = subprocess.run([sys.executable, os.path.join(REF, "postprocess_cobs.py"), "-n", str(keep)], input=raw,
                                  check=True, stdout=subprocess.PIPE).stdout
        else:
            post = filters.postprocess_text(raw.decode(), keep).encode()
        p = os.path.join(mdir, f"{b}____q.gz")
        with gzip.open(p, "wb") as f:
            f.write(post)
        files.append(p)
        texts[b] = post.decode()
    if HAVE_REF:
        env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oracle", "xopen_shim"))
        fa = subprocess.run([sys.executable, os.path.join(REF, "filter_queries.py"), "-n", str(keep), "-q", qfa] + files,
                            check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env).stdout.decode()
    else:
        from phylign_b200.fasta import read_fastx
        qs = read_fastx(qfa)
        bs = [(b, [(h.split(" ")[0], [(n.split("_")[1], s) for n, s in hits])
                   for h, _, hits in filters.parse_cobs_text(texts[b])]) for b in batches]
        d = {}
        for q, s in qs:
            d[q] = s
        fa = filters.merge_running(list(d.items()), bs, keep)
    return texts, fa


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rnd = random.Random(seed)
    oracle.build()
    gpus = n_gpus()
    ok = 0
    for case in range(n_cases):
        td = tempfile.mkdtemp(prefix="phy_fuzzpipe_")
        try:
            os.makedirs(os.path.join(td, "cobs"))
            cpu_dir = os.path.join(td, "cobs_cpu")
            os.makedirs(cpu_dir)
            glen = rnd.choice([600, 1500, 2500])
            root = "".join(rnd.choice("ACGT") for _ in range(glen))
            batches = []
            for bi in range(rnd.randrange(2, 7)):
                b = f"{rnd.choice(['aa', 'bb', 'zz', 'mm'])}{rnd.randrange(100):02d}_sp__{bi:02d}"
                n_docs = rnd.choice([3, 9, 60, 130, 300, 700, 1100, 2100])
                docs, names = [], []
                prefixes = sorted(rnd.sample(range(10 ** 6), n_docs))
                for d in range(n_docs):
                    if d < 4 or rnd.random() < 6.0 / n_docs:
                        s = list(root)
                        for _ in range(rnd.randrange(0, glen // 15)):
                            s[rnd.randrange(glen)] = rnd.choice("ACGT")
                        docs.append("".join(s).encode())
                    else:
                        docs.append(b"")
                    names.append(f"{prefixes[d]:06d}_ACC{bi}x{rnd.randrange(10 ** 6):06d}x{d}")
                oi = oracle.OracleIndex.construct(docs, names, num_hashes=rnd.choice([1, 1, 1, 2]),
                                                  signature_size_override=rnd.choice([None, 4099, 20011]))
                cpu_path = os.path.join(cpu_dir, f"{b}.cobs_classic")
                oi.write(cpu_path)
                if rnd.random() < 0.5:      # half of the batches only exist as .xz for the drop-in, like the real database
                    with open(os.path.join(td, "cobs", f"{b}.cobs_classic.xz"), "wb") as f:
                        subprocess.run(["xz", "-1", "-c", cpu_path], check=True, stdout=f)
                else:
                    shutil.copy(cpu_path, os.path.join(td, "cobs", f"{b}.cobs_classic"))
                batches.append(b)
            # queries
            plain = rnd.random() < 0.5
            recs = []
            for j in range(rnd.randrange(1, 40)):
                ln = min(glen, rnd.choice([20, 31, 40, 150, 300, 1100, rnd.randrange(1, glen)]))
                a = rnd.randrange(0, glen - ln + 1)
                s = root[a:a + ln] if rnd.random() < 0.8 else "".join(rnd.choice("ACGT") for _ in range(ln))
                name = f"q{j}" if plain or rnd.random() < 0.9 else f"q{rnd.randrange(max(1, j))}"   # duplicate names
                recs.append((name, s))
            qfa = os.path.join(td, "q.fa")
            with open(qfa, "w") as f:
                for name, s in recs:
                    if plain:
                        f.write(f">{name}\n{s}\n")
                    else:
                        head = name + (" some comment" if rnd.random() < 0.3 else "")
                        w = rnd.choice([len(s), 60, 17])
                        f.write(">" + head + "\n" + "\n".join(s[i:i + w] for i in range(0, len(s), w)) + "\n")
                        if rnd.random() < 0.1:
                            f.write("\n")
            with open(os.path.join(td, "batches.txt"), "w") as f:
                f.write("\n".join(batches) + "\n")
            thr = rnd.choice([0.3, 0.5, 0.7, 0.7, 0.9])
            keep = rnd.choice([1, 2, 5, 100])
            # (a) CPU pipeline on the decompressed copies
            want_txt, want_fa = cpu_pipeline(td, cpu_dir, batches, qfa, thr, keep)
            # (b) the drop-in
            cmd = [sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", os.path.join(td, "cobs"), "--batches",
                   os.path.join(td, "batches.txt"), "-q", qfa, "--qfile", "q", "--match-dir", os.path.join(td, "03"),
                   "--filter-out", os.path.join(td, "04", "q.fa"), "-t", str(thr), "-n", str(keep)]
            opts = []
            if rnd.random() < 0.5:
                opts += ["--query-block-bases", str(rnd.choice([200, 2000, 20000]))]
            if rnd.random() < 0.5:
                opts += ["--round-bytes", str(rnd.choice([3_000_000, 30_000_000]))]
            if rnd.random() < 0.3:
                opts += ["--no-overlap-rounds"]
            if gpus > 1 and rnd.random() < 0.6:
                opts += ["--gpus", str(rnd.choice(range(2, min(gpus, 4) + 1)))]
            if rnd.random() < 0.3:
                opts += ["--decompression-dir", os.path.join(td, "dec"), "--keep-cobs-indexes"]
            r = subprocess.run(cmd + opts, capture_output=True, text=True, env=ENV, cwd=ROOT, timeout=600)
            if r.returncode != 0 and "larger than the per-GPU HBM budget" in r.stderr:
                opts = [o for o in opts if o not in ("--round-bytes", "3000000", "30000000")]
                r = subprocess.run(cmd + opts, capture_output=True, text=True, env=ENV, cwd=ROOT, timeout=600)
            assert r.returncode == 0, (case, opts, r.stderr[-2000:])
            if rnd.random() < 0.4:           # --resume after losing some outputs
                lost = rnd.sample(batches, rnd.randrange(1, len(batches) + 1))
                for b in lost:
                    os.unlink(os.path.join(td, "03", f"{b}____q.gz"))
                os.unlink(os.path.join(td, "04", "q.fa"))
                r = subprocess.run(cmd + opts + ["--resume"], capture_output=True, text=True, env=ENV, cwd=ROOT, timeout=600)
                assert r.returncode == 0, (case, opts, "resume", r.stderr[-2000:])
            for b in batches:
                got = gzip.open(os.path.join(td, "03", f"{b}____q.gz"), "rt").read()
                assert canon_match_text(got) == canon_match_text(want_txt[b]), (case, b, opts)
            got_fa = open(os.path.join(td, "04", "q.fa")).read()
            assert got_fa == want_fa, (case, opts, plain)
            assert not [f for f in os.listdir(os.path.join(td, "04")) if f != "q.fa"], "stray part / tmp files"
            ok += 1
            if (case + 1) % 5 == 0:
                print(f"case {case + 1}: ok so far ({'reference scripts' if HAVE_REF else 'filters.py port'})", flush=True)
        finally:
            shutil.rmtree(td, ignore_errors=True)
    print(f"fuzz pipeline: {ok} random databases/query sets (seed {seed}) identical to the CPU pipeline "
          f"({'unmodified postprocess_cobs.py + filter_queries.py' if HAVE_REF else 'oracle/filters.py port'}); "
          f"{gpus} GPU(s) visible")


if __name__ == "__main__":
    main()
