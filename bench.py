#!/usr/bin/env python3
"""bench.py -- COBS-match throughput of the B200 path, and of the CPU reference pipeline beside it.

Workload (BASELINE.json configs[2]): 100k synthetic 1 kbp reads vs 64 synthetic batch
indexes (4000 docs x 1 Mbp genomes each, k=31, 1 hash, fpr 0.3 -> ~1.43 GB per index)
resident in HBM, threshold 0.7, top-N 100 + ties, per-query merge over all indexes.
A step = one pass of all reads over all indexes.  N GPUs: the 64 indexes are placed on the
ranks (strong scaling, fixed total work), queries replicated, per-GPU candidate lists gathered
with NCCL and merged inside the step.

One JSON line (rank 0):
  value          device-timed bases/s, queries resident in HBM
  e2e            the same through the C ABI with host buffers (H2D + D2H inside the timed region)
  e2e_files      files in -> files out: `python -m phylign_b200.cli match-db` on the same workload
                 written to /dev/shm (.cobs_classic files + FASTA) producing 03_match/*.gz and
                 04_filter/*.fa, with the wall-clock breakdown; also for 150-bp reads
  roofline       HBM roofline of the gather kernel;  cpu_baseline: the oracle on the host cores
  result_digest  sha256 over canonicalised (units, hits, merged): identical for every N and for
                 pruning on/off (asserted here; the run fails otherwise)
  secondary      150-bp read class; at --gpus 8 also the full 661k-shaped database (configs[3])

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

--impl reference times the reference's own CPU pipeline for the same workload on a bounded sample:
the CPU restatement of `cobs query` (oracle/, -O3 -march=native, all host cores) and, for e2e_files,
`cobs_oracle query | postprocess_cobs.py -n 100 | gzip --fast` per batch followed by
`filter_queries.py` (the unmodified reference scripts from baseline/_ref/scripts when present).
That process never loads libphylign_cuda.so: its inputs are written by a separate set-up process.
"""
import argparse
import hashlib
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "COBS-match query bases/s over whole index"
UNIT = "bases/s"
THRESHOLD, TOP_N = 0.7, 100
READS_SEED, RANDOM_Q8, ERR_Q16 = 3, 51, 655
SHORT_READ_LEN = 150


def workload(args):
    """Batches (indexes) + reads of the selected workload.

    reads1k (default, BASELINE configs[2]): --indexes equal synthetic batches.
    db661k (BASELINE configs[3]): the 305 batches of the 661k database, document counts and
    decompressed sizes from the reference's data files (phylign_b200/data/db_shape.tsv),
    content synthetic; --db-scale shrinks every signature_size (1.0 = the real 1.06 TB)."""
    ratio = -1.0 / math.log(1.0 - 0.3)                      # cobs classic-construct: fpr 0.3, 1 hash
    batches = []
    if args.workload == "db661k":
        for line in open(os.path.join(ROOT, "phylign_b200", "data", "db_shape.tsv")):
            if line.startswith("#"):
                continue
            name, nbytes, docs = line.split("\t")
            docs = int(docs)
            sig = max(4096, int(int(nbytes) / ((docs + 7) // 8) * args.db_scale))
            batches.append(dict(name=name, n_docs=docs, signature_size=sig,
                                genome_len=max(1200, int(sig / ratio)), seed=1000 + len(batches)))
    else:
        sig = int(math.ceil((args.genome_len - 30) * ratio))
        for i in range(args.indexes):
            batches.append(dict(name=batch_name(i), n_docs=args.docs, signature_size=sig,
                                genome_len=args.genome_len, seed=1000 + i))
    w = dict(batches=batches, n_indexes=len(batches), n_docs=max(b["n_docs"] for b in batches),
             n_reads=args.reads, read_len=args.read_len, name=args.workload)
    w["signature_size"] = batches[0]["signature_size"]
    w["row_bytes_per_kmer"] = sum((b["n_docs"] + 7) // 8 for b in batches)
    w["kmers_per_read"] = max(w["read_len"] - 30, 0)
    w["bases"] = w["n_reads"] * w["read_len"]
    # algorithmic bytes = sum_q K_q * sum_b h_b * ceil(D_b/8)
    w["alg_bytes"] = w["n_reads"] * w["kmers_per_read"] * w["row_bytes_per_kmer"]
    w["kmer_docs"] = w["n_reads"] * w["kmers_per_read"] * sum(b["n_docs"] for b in batches)
    return w


def with_read_len(w, read_len, n_reads=None):
    """The same batches queried with reads of another length."""
    v = dict(w)
    v["read_len"], v["n_reads"] = read_len, n_reads or w["n_reads"]
    v["kmers_per_read"] = max(read_len - 30, 0)
    v["bases"] = v["n_reads"] * read_len
    v["alg_bytes"] = v["n_reads"] * v["kmers_per_read"] * v["row_bytes_per_kmer"]
    v["kmer_docs"] = v["n_reads"] * v["kmers_per_read"] * sum(b["n_docs"] for b in v["batches"])
    return v


def spec_kwargs(b, w=None):
    if isinstance(b, int):                                   # index number of the reads1k workload
        b = w["batches"][b]
    return dict(seed=b["seed"], n_docs=b["n_docs"], genome_len=b["genome_len"], clade_size=32,
                clade_sub_q16=328, doc_sub_q16=328)


def batch_name(i):
    return f"synth_species_{i:03d}__01"


def place(w, world, budget):
    """LPT placement of the batches on the ranks (phylign_b200/sharding.py); one resident round."""
    from phylign_b200 import sharding
    bl = [sharding.Batch(b["name"], b["n_docs"], b["signature_size"]) for b in w["batches"]]
    plan = sharding.assign(bl, world, budget)
    if len(plan.rounds) != 1:
        raise SystemExit(f"workload needs {len(plan.rounds)} resident rounds on {world} GPU(s): "
                         "use more GPUs or --db-scale")
    by_name = {b["name"]: b for b in w["batches"]}
    return [[by_name[x.name] for x in plan.batches_of(r)] for r in range(world)], plan.imbalance


def config_dict(w):
    """The workload, identically worded in both arms (run-dependent facts go under "run")."""
    if w["name"] == "db661k":
        what = (f"{w['n_indexes']} synthetic COBS classic indexes shaped like the 661k database "
                f"(docs and sizes per batch from data/decompressed_indexes_sizes.txt + 661k_batches.txt, "
                f"{w['row_bytes_per_kmer']} B gathered per k-mer) (BASELINE.json configs[3])")
    else:
        what = (f"{w['n_indexes']} synthetic COBS classic indexes ({w['n_docs']} docs x "
                f"{w['batches'][0]['genome_len']} bp, k=31, h=1, fpr=0.3) (BASELINE.json configs[2])")
    return {"workload": f"{w['n_reads']} synthetic {w['read_len']} bp reads vs {what} resident in HBM, "
                        f"-t {THRESHOLD}, top-{TOP_N}+ties, cross-index merge",
            "n_reads": w["n_reads"], "read_len": w["read_len"], "n_indexes": w["n_indexes"],
            "docs_per_index": w["n_docs"], "threshold": THRESHOLD, "top_n": TOP_N,
            "algorithmic_bytes_per_step": w["alg_bytes"], "kmer_docs_per_step": w["kmer_docs"],
            "l2_hygiene": "inputs (index shard) exceed the 126 MB L2; no flush needed"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def shm_dir(prefix):
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.mkdtemp(prefix=prefix, dir=base)


def write_fasta(path, raw: bytes, n_reads: int, read_len: int, name="read"):
    """>read_0000000\\nSEQ\\n ... built with numpy (no per-record Python loop)."""
    names = np.char.add(f">{name}_", np.char.zfill(np.arange(n_reads).astype("U7"), 7)).astype("S")
    w = names.dtype.itemsize
    rec = np.empty((n_reads, w + 1 + read_len + 1), dtype=np.uint8)
    rec[:, :w] = names.view(np.uint8).reshape(n_reads, w)
    rec[:, w] = ord("\n")
    rec[:, w + 1:w + 1 + read_len] = np.frombuffer(raw, dtype=np.uint8, count=n_reads * read_len).reshape(n_reads, read_len)
    rec[:, -1] = ord("\n")
    with open(path, "wb") as f:
        f.write(rec.tobytes())


# =============================================================================== CPU arm
def reference_scripts():
    """(postprocess_cobs.py, filter_queries.py, kind): the UNMODIFIED reference scripts when
    __graft_entry__.build() installed them under baseline/_ref/scripts (git-ignored, travels to the
    GPU box), else None -> the restatement in oracle/filters.py is timed and labelled "port"."""
    d = os.path.join(ROOT, "baseline", "_ref", "scripts")
    pp, fq = os.path.join(d, "postprocess_cobs.py"), os.path.join(d, "filter_queries.py")
    if os.path.exists(pp) and os.path.exists(fq):
        return pp, fq, "reference"
    return None, None, "port"


def find_cobs():
    """An executable `cobs` on PATH (never the repo's own scripts/cobs front end), else None."""
    for d in os.environ.get("PATH", "").split(os.pathsep):
        p = os.path.join(d, "cobs")
        if os.path.isfile(p) and os.access(p, os.X_OK) and os.path.realpath(p) != os.path.realpath(
                os.path.join(ROOT, "scripts", "cobs")):
            try:
                head = open(p, "rb").read(256)
            except OSError:
                continue
            if b"phylign_b200" in head:
                continue
            return p
    return None


def time_oracle_layouts(oidx, reads, threads):
    """{(simd, layout): seconds} for one pass of `reads` over ONE index."""
    import oracle
    out = {}
    for avx in (False, True):
        simd = oracle.set_simd(avx)
        if avx and simd != "avx2":
            continue
        for mode, name in ((0, "queries serial, 128-doc slices over threads (cobs layout)"),
                           (1, "queries over threads")):
            t0 = time.perf_counter()
            oidx.query_batch(reads, THRESHOLD, threads=threads, mode=mode)
            out[(simd, name)] = time.perf_counter() - t0
    return out


def cpu_pipeline_files(workdir, index_path, batch, sample_fa, n_sample, read_len, n_indexes, cores):
    """The reference's file pipeline on the sample (Snakefile:416-428 then :513-520):
        cobs query -t 0.7 -T cores -i INDEX -f reads.fa | postprocess_cobs.py -n 100 | gzip --fast > match.gz
    for ONE batch (timed; the database is n_indexes such jobs), then filter_queries.py over n_indexes
    match files (the one produced, under n_indexes batch names) -> 04_filter FASTA."""
    import oracle
    pp, fq, kind = reference_scripts()
    cobs = find_cobs()
    cobs_cmd = [cobs, "query", "--load-complete"] if cobs else [oracle.cli_path(), "query", "--load-complete"]
    mdir = os.path.join(workdir, "ref_03_match")
    os.makedirs(mdir, exist_ok=True)
    match0 = os.path.join(mdir, f"{batch}____reads.gz")
    q = lambda s: "'" + s.replace("'", "'\\''") + "'"
    if pp:
        post = f"{q(sys.executable)} {q(pp)} -n {TOP_N}"
    else:
        post = (f"{q(sys.executable)} -c \"import sys; sys.path.insert(0, {ROOT!r}); from oracle import filters; "
                f"sys.stdout.write(filters.postprocess_text(sys.stdin.read(), {TOP_N}))\"")
    pipe = (f"set -euo pipefail; {' '.join(q(c) for c in cobs_cmd)} -t {THRESHOLD} -T {cores} -i {q(index_path)} "
            f"-f {q(sample_fa)} | {post} | gzip --fast > {q(match0)}")
    t0 = time.perf_counter()
    subprocess.run(["bash", "-c", pipe], check=True)
    t_batch = time.perf_counter() - t0
    # stage by stage (explains the pipe; the stages overlap inside it)
    t0 = time.perf_counter()
    raw_txt = subprocess.run(cobs_cmd + ["-t", str(THRESHOLD), "-T", str(cores), "-i", index_path, "-f", sample_fa],
                             check=True, stdout=subprocess.PIPE).stdout
    t_cobs = time.perf_counter() - t0
    t0 = time.perf_counter()
    post_txt = subprocess.run(["bash", "-c", post], input=raw_txt, check=True, stdout=subprocess.PIPE).stdout
    t_post = time.perf_counter() - t0
    t0 = time.perf_counter()
    subprocess.run(["gzip", "--fast", "-c"], input=post_txt, check=True, stdout=subprocess.DEVNULL)
    t_gzip = time.perf_counter() - t0
    # translate_matches over all batches: the same content under n_indexes batch names
    files = [match0]
    for i in range(1, n_indexes):
        p = os.path.join(mdir, f"{batch_name(i)}____reads.gz")
        if os.path.abspath(p) != os.path.abspath(match0):
            shutil.copyfile(match0, p)
            files.append(p)
    out_fa = os.path.join(workdir, "ref_04_filter.fa")
    env = dict(os.environ)
    if fq:
        shim = os.path.join(ROOT, "oracle", "xopen_shim")
        env["PYTHONPATH"] = shim + os.pathsep + env.get("PYTHONPATH", "")
        cmd = [sys.executable, fq, "-n", str(TOP_N), "-q", sample_fa] + files
    else:
        cmd = [sys.executable, "-c",
               "import sys, gzip; sys.path.insert(0, %r); from oracle import filters\n"
               "qs=[]; name=None\n"
               "for l in open(sys.argv[1]):\n"
               "    l=l.rstrip()\n"
               "    if l.startswith('>'): name=l[1:].split(' ')[0]\n"
               "    elif name is not None: qs.append((name,l)); name=None\n"
               "bs=[]\n"
               "for fn in sys.argv[2:]:\n"
               "    b=fn.split('/')[-1].split('____')[0]\n"
               "    bs.append((b,[(h.split(' ')[0],[(n.split('_')[1],s) for n,s in hits]) for h,_,hits in "
               "filters.parse_cobs_text(gzip.open(fn,'rt').read())]))\n"
               "sys.stdout.write(filters.merge_running(qs,bs,%d))\n" % (ROOT, TOP_N), sample_fa] + files
    t0 = time.perf_counter()
    with open(out_fa, "wb") as fo:
        subprocess.run(cmd, check=True, stdout=fo, stderr=subprocess.DEVNULL, env=env)
    t_filter = time.perf_counter() - t0
    total = t_batch * n_indexes + t_filter
    return {"value": n_sample * read_len / total, "unit": UNIT,
            "wall_s_for_sample": total,
            "breakdown_s": {"match_pipeline_one_batch": round(t_batch, 3),
                            "match_pipeline_all_batches_extrapolated": round(t_batch * n_indexes, 3),
                            "stage_cobs_query_alone": round(t_cobs, 3), "stage_postprocess_alone": round(t_post, 3),
                            "stage_gzip_fast_alone": round(t_gzip, 3),
                            "filter_queries_all_batches": round(t_filter, 3)},
            "sample": f"{n_sample} reads: `{'cobs' if cobs else 'cobs_oracle'} query -T {cores} | postprocess_cobs.py -n {TOP_N} "
                      f"| gzip --fast` timed on 1 batch and counted x{n_indexes} (one job per batch, all cores each), then "
                      f"filter_queries.py over {len(files)} match files (measured)",
            "scripts": kind + (" (unmodified postprocess_cobs.py / filter_queries.py from baseline/_ref/scripts)"
                               if kind == "reference" else " (oracle/filters.py restatement; reference scripts not installed)"),
            "cobs_binary": cobs or "absent -> oracle port",
            "outputs": {"match_file_bytes": os.path.getsize(match0), "filter_fasta_bytes": os.path.getsize(out_fa)}}


def cmd_reference_setup(args, w):
    """Set-up process of the CPU arm: index 0 and the sample reads as files.  Uses the GPU library
    only to synthesise them quickly (falls back to the oracle's own generator for small shapes)."""
    n_sample = min(w["n_reads"], args.cpu_sample_reads)
    specs_kw = [spec_kwargs(b) for b in w["batches"]]
    w0 = w["batches"][0]
    os.makedirs(os.path.join(args.workdir, "cobs"), exist_ok=True)
    ipath = os.path.join(args.workdir, "cobs", f"{w0['name']}.cobs_classic")
    try:
        from phylign_b200 import _lib
        from phylign_b200.matcher import Matcher
        m = Matcher(int(os.environ.get("LOCAL_RANK", 0)))
        i0 = m.add_synth_index(w0["name"], _lib.SynthSpec(**specs_kw[0]), w0["signature_size"])
        m.write_index_file(i0, ipath)
        raw = m.synth_reads([_lib.SynthSpec(**k) for k in specs_kw], READS_SEED, 0, n_sample, w["read_len"],
                            RANDOM_Q8, ERR_Q16)
        raw150 = m.synth_reads([_lib.SynthSpec(**k) for k in specs_kw], READS_SEED, 0, n_sample, SHORT_READ_LEN,
                               RANDOM_Q8, ERR_Q16)
        m.close()
        built = "index 0 and the reads synthesised on the GPU by a separate set-up process"
    except Exception as e:  # no GPU: build the (small) workload with the oracle itself
        if w0["n_docs"] * w0["genome_len"] > 5e7:
            print(json.dumps({"error": f"cannot synthesise the index without a GPU: {e}"}))
            return 1
        import oracle
        ospecs = [oracle.SynthSpec(**k) for k in specs_kw]
        names = [f"r{d:06d}_SYN{d:06d}" for d in range(w0["n_docs"])]
        oidx = oracle.OracleIndex.construct([oracle.synth_genome(ospecs[0], d) for d in range(w0["n_docs"])],
                                            names, signature_size_override=w0["signature_size"])
        oidx.write(ipath)
        raw = b"".join(oracle.synth_read(ospecs, READS_SEED, r, w["read_len"], RANDOM_Q8, ERR_Q16)
                       for r in range(n_sample))
        raw150 = b"".join(oracle.synth_read(ospecs, READS_SEED, r, SHORT_READ_LEN, RANDOM_Q8, ERR_Q16)
                          for r in range(n_sample))
        built = "index 0 and the reads built by the oracle on the CPU (no GPU visible)"
    write_fasta(os.path.join(args.workdir, "reads.fa"), raw, n_sample, w["read_len"])
    write_fasta(os.path.join(args.workdir, "reads150.fa"), raw150, n_sample, SHORT_READ_LEN)
    print(json.dumps({"built": built, "index": ipath, "n_sample": n_sample}))
    return 0


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    import oracle
    native = oracle.build_native()
    oracle.build()
    cores = host_cores()
    workdir = shm_dir("phylign_ref_")
    try:
        argv = [sys.executable, os.path.abspath(__file__), "--impl", "reference-setup", "--workdir", workdir]
        for k in ("reads", "read_len", "indexes", "docs", "genome_len", "cpu_sample_reads", "workload", "db_scale"):
            argv += ["--" + k.replace("_", "-"), str(getattr(args, k))]
        r = subprocess.run(argv, capture_output=True, text=True)
        info = {}
        try:
            info = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            pass
        if r.returncode != 0 or "error" in info:
            print(json.dumps({"impl": "reference", "unavailable": (info.get("error") or r.stderr[-300:]).replace("\n", " ")}))
            return
        assert "phylign_b200._lib" not in sys.modules and "phylign_b200.matcher" not in sys.modules
        n_sample, L = info["n_sample"], w["read_len"]
        oidx = oracle.OracleIndex.read(info["index"])
        txt = open(os.path.join(workdir, "reads.fa"), "rb").read().split(b"\n")
        reads = [txt[2 * i + 1] for i in range(n_sample)]
        probe = reads[:max(1, n_sample // 8)]
        lay = time_oracle_layouts(oidx, probe, cores)                       # pick the fastest kernel + layout
        best = min(lay, key=lay.get)
        oracle.set_simd(best[0] == "avx2")
        mode = 0 if best[1].startswith("queries serial") else 1
        for _ in range(args.warmup):
            oidx.query_batch(probe, THRESHOLD, threads=cores, mode=mode)
        per_step = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            oidx.query_batch(reads, THRESHOLD, threads=cores, mode=mode)
            per_step.append(time.perf_counter() - t0)
        dt = float(np.mean(per_step))
        kind = "port"
        cobs = find_cobs()
        if cobs:     # a real cobs binary is on PATH: it IS the reference -- time it instead of the port
            cmd = [cobs, "query", "--load-complete", "-t", str(THRESHOLD), "-T", str(cores), "-i", info["index"],
                   "-f", os.path.join(workdir, "reads.fa")]
            try:
                subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)   # warm-up
                per_step = []
                for _ in range(args.steps):
                    t0 = time.perf_counter()
                    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                    per_step.append(time.perf_counter() - t0)
                dt = float(np.mean(per_step))
                kind = "reference"
            except Exception:
                cobs = None
        # one index timed; the whole database is n_indexes such passes (independent batches)
        value = n_sample * L / (dt * w["n_indexes"])
        sample = (f"{n_sample} of the {w['n_reads']} reads against 1 of the {w['n_indexes']} indexes per step "
                  f"({dt:.2f} s, min {min(per_step):.2f} max {max(per_step):.2f}), extrapolated x{w['n_indexes']} indexes; "
                  + (f"`{cobs} query --load-complete -T {cores}` (the cobs binary found on PATH; index load included in "
                     f"every call, as in the Snakemake rule); " if kind == "reference" else
                     f"CPU restatement of cobs 0.2.1 classic query (oracle port, NOT the cobs binary), "
                     f"{'-O3 -march=native' if native else '-O3 -march=x86-64-v2'}, kernel {best[0]}, layout '{best[1]}'; ")
                  + f"{info['built']}")
        files = None
        if not args.no_e2e_files:
            try:
                files = cpu_pipeline_files(workdir, info["index"], w["batches"][0]["name"],
                                           os.path.join(workdir, "reads.fa"), n_sample, L, w["n_indexes"], cores)
            except Exception as e:
                files = {"value": None, "error": f"{type(e).__name__}: {e}"}
        # the short-read class beside it (the same sample size, 150-bp reads)
        secondary = {}
        try:
            txt150 = open(os.path.join(workdir, "reads150.fa"), "rb").read().split(b"\n")
            reads150 = [txt150[2 * i + 1] for i in range(n_sample)]
            oidx.query_batch(reads150[:max(1, n_sample // 8)], THRESHOLD, threads=cores, mode=mode)
            t0 = time.perf_counter()
            for _ in range(3):
                oidx.query_batch(reads150, THRESHOLD, threads=cores, mode=mode)
            dt150 = (time.perf_counter() - t0) / 3
            w150 = with_read_len(w, SHORT_READ_LEN)
            secondary["reads150"] = {"workload": config_dict(w150)["workload"],
                                     "value": n_sample * SHORT_READ_LEN / (dt150 * w["n_indexes"]), "unit": UNIT,
                                     "sample": f"{n_sample} reads x 1 index ({dt150:.2f} s), x{w['n_indexes']}"}
            if not args.no_e2e_files:
                secondary["reads150"]["e2e_files"] = cpu_pipeline_files(
                    os.path.join(workdir, "p150"), info["index"], w["batches"][0]["name"],
                    os.path.join(workdir, "reads150.fa"), n_sample, SHORT_READ_LEN, w["n_indexes"], cores)
        except Exception as e:
            secondary["reads150"] = {"value": None, "error": f"{type(e).__name__}: {e}"}
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": config_dict(w),
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                                 "variants_s_on_probe": {f"{k[0]} | {k[1]}": round(v, 3) for k, v in lay.items()},
                                 "cobs_on_path": find_cobs()},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "e2e_files": files,
                "secondary": secondary,
                "native_so_policy": "this process loaded only oracle/_build; inputs came from a set-up subprocess"}
        print(json.dumps(line))
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


# =============================================================================== GPU arm
class Dist:
    """gloo control plane under torchrun (barrier, max, gather of small objects); no-ops at N=1."""

    def __init__(self, rank, world, local_rank):
        self.rank, self.world, self.d = rank, world, None
        if world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local_rank)
            dist.init_process_group("gloo", rank=rank, world_size=world)
            self.d, self.torch = dist, torch

    def barrier(self):
        if self.d is not None:
            self.d.barrier()

    def max(self, x):
        if self.d is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64)
        self.d.all_reduce(t, op=self.d.ReduceOp.MAX)
        return float(t[0])

    def sum(self, x):
        if self.d is None:
            return x
        t = self.torch.tensor([float(x)], dtype=self.torch.float64)
        self.d.all_reduce(t, op=self.d.ReduceOp.SUM)
        return float(t[0])

    def gather(self, obj):
        """[obj of every rank] on rank 0 (None elsewhere)."""
        if self.d is None:
            return [obj]
        out = [None] * self.world if self.rank == 0 else None
        self.d.gather_object(obj, out, dst=0)
        return out

    def bcast(self, obj):
        if self.d is None:
            return obj
        box = [obj]
        self.d.broadcast_object_list(box, src=0)
        return box[0]


def multi_gpu_options(m, args):
    """N > 1: every rank finalises and downloads its own slice of the queries' merged lists and uploads
    1/N of the bases (all-gather over NVLink) -- what `match-db --gpus N` does too."""
    m.set_option("merge_mode", 0 if args.merge_on_rank0 else 1)
    m.set_option("shard_query_upload", 0 if args.merge_on_rank0 else 1)


def full_merged(dist, m, moffs, mcands):
    """(offs, cands) of ALL queries on rank 0 (None elsewhere): the ranks' slices joined in query order
    when the merge is query-sharded.  Control-plane traffic, outside every timed region."""
    if dist.world == 1:
        return np.asarray(moffs), np.asarray(mcands)
    lo, hi = m.merged_range()
    o = np.asarray(moffs).astype(np.int64)
    parts = dist.gather((lo, hi, o[lo:hi + 1] - o[lo], np.array(mcands)))
    if dist.rank != 0:
        return None, None
    parts = sorted((p for p in parts if p[1] > p[0]), key=lambda p: p[0])
    nq = len(o) - 1
    offs = np.zeros(nq + 1, dtype=np.uint64)
    cands, run = [], 0
    for lo, hi, po, pc in parts:
        offs[lo:hi + 1] = (po + run).astype(np.uint64)
        run += int(po[-1])
        offs[hi:] = run
        cands.append(pc)
    return offs, (np.concatenate(cands) if cands else np.array(mcands)[:0])


def merged_struct(offs, cands):
    """A phy_merged over host arrays (for the library's formatters); keep the arrays alive while it is used."""
    import ctypes as C
    from phylign_b200 import _lib
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    cands = np.ascontiguousarray(cands)
    st = _lib.Merged(len(offs) - 1, C.cast(offs.ctypes.data, C.POINTER(C.c_uint64)),
                     C.cast(cands.ctypes.data, C.POINTER(_lib.Cand)), 0)
    return st, (offs, cands)


def unit_hits_in_order(res):
    """hits re-ordered so that unit i's hits sit at [starts[i], starts[i]+n_kept[i]) (units are sorted
    by (index, query); their hit blocks are not, they were claimed with an atomic cursor)."""
    u = res.units
    kept = u["n_kept"].astype(np.int64)
    starts = np.concatenate(([0], np.cumsum(kept)[:-1])) if len(u) else np.zeros(0, np.int64)
    total = int(kept.sum())
    pos = np.repeat(u["offset"].astype(np.int64) - starts, kept) + np.arange(total, dtype=np.int64)
    return starts, (res.hits[pos] if total else res.hits[:0])


def local_digests(m, res):
    """{batch_rank: sha256} over (query, n_pass, n_kept) of the index's units + their hits (doc, score)."""
    u = res.units
    starts, hits = unit_hits_in_order(res)
    out = {}
    for idx, ix in m.indexes.items():
        lo, hi = np.searchsorted(u["index"], idx, "left"), np.searchsorted(u["index"], idx, "right")
        h = hashlib.sha256()
        part = u[lo:hi]
        h.update(np.ascontiguousarray(part["query"]).tobytes())
        h.update(np.ascontiguousarray(part["n_pass"]).tobytes())
        h.update(np.ascontiguousarray(part["n_kept"]).tobytes())
        if hi > lo:
            h0 = int(starts[lo])
            h1 = int(starts[hi - 1] + part["n_kept"][-1])
            h.update(np.ascontiguousarray(hits[h0:h1]).tobytes())
        out[int(ix.batch_rank)] = h.hexdigest()
    return out


def result_digest(dist, m, res, moffs, mcands):
    """sha256 over the per-batch digests in batch-rank order + the merged lists: independent of the
    number of GPUs and of the placement (idx ids are local, batch ranks are global)."""
    parts = dist.gather(local_digests(m, res))
    moffs, mcands = full_merged(dist, m, moffs, mcands)
    if dist.rank != 0:
        return None
    allb = {}
    for p in parts:
        allb.update(p)
    h = hashlib.sha256()
    for br in sorted(allb):
        h.update(f"{br}:{allb[br]};".encode())
    h.update(np.ascontiguousarray(moffs).tobytes())
    h.update(np.ascontiguousarray(mcands).tobytes())
    return h.hexdigest()


def timed_steps(dist, m, steps, warmup, rank, local_rank, sample_clocks=False):
    """W untimed + K timed device-resident steps (barrier + sync on both sides, CUDA events on the
    library's stream, max over ranks).  Returns a dict."""
    for _ in range(warmup):
        m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
    sampler = ClockSampler(local_rank) if (rank == 0 and sample_clocks) else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    m.sync()
    dist.barrier()
    launches, gather_ms, phases, gathered = 0, [], [], 0
    m.timer_start()
    for _ in range(steps):
        m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
        gathered = m.gathered_bytes()
        ph = m.phase_ms()
        gather_ms.append(ph[1])
        phases.append(ph[:3])
        launches += int(ph[3])
    dev_ms = m.timer_stop()
    m.sync()
    dist.barrier()
    clocks = sampler.stop() if sampler else None
    return {"ms_per_step": dist.max(dev_ms) / steps, "gather_ms": float(np.mean(gather_ms)),
            "phases": [round(float(x), 3) for x in np.mean(phases, axis=0)], "gathered": int(gathered),
            "launches": launches, "clocks": clocks}


def pinned_queries(raw, offs):
    from phylign_b200.matcher import PinnedBuffer
    pr, po = PinnedBuffer(len(raw)), PinnedBuffer(offs.nbytes)
    pr.array[:] = np.frombuffer(raw, dtype=np.uint8)
    po.array[:] = offs.view(np.uint8)
    return pr, po


def e2e_steps(dist, m, pin_raw, pin_offs, raw_len, offs_bytes, steps, warmup, sharded_upload=False):
    for _ in range(max(2, warmup)):   # warm-up: the pinned result pool reaches its steady state after 2 passes
        m.set_queries_raw(pin_raw, pin_offs)
        m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
        res = m.fetch()               # held like in the timed loop, so the pool ends up with both
        moffs, mcands = m.merged()    # generations of result blocks before timing starts
    m.sync()
    dist.barrier()
    h2d = d2h = 0
    parts = np.zeros(4)
    t0 = time.perf_counter()
    for _ in range(steps):
        ta = time.perf_counter()
        m.set_queries_raw(pin_raw, pin_offs)               # H2D of the step's inputs (pinned host buffers)
        tb = time.perf_counter()
        m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
        tc = time.perf_counter()
        res = m.fetch()                                    # D2H: per-(query,index) hit lists (03_match content)
        td = time.perf_counter()
        moffs, mcands = m.merged()                         # D2H: merged top-N lists (04_filter content)
        te = time.perf_counter()
        parts += [tb - ta, tc - tb, td - tc, te - td]
        h2d = (raw_len + dist.world - 1) // dist.world + offs_bytes if sharded_upload else raw_len + offs_bytes
        d2h = res.d2h_bytes + moffs.nbytes + mcands.nbytes
    m.sync()
    e2e_s = dist.max((time.perf_counter() - t0) / steps)
    dist.barrier()
    return e2e_s, h2d, d2h, [round(float(x) * 1e3 / steps, 2) for x in parts], res, moffs, mcands


def write_index_files(m, ids, outdir, n_bufs=4):
    """The resident indexes as {batch}.cobs_classic files (set-up of the files-in -> files-out run):
    downloads go through a few page-locked buffers, the file writes run on threads behind them."""
    from concurrent.futures import ThreadPoolExecutor
    from phylign_b200.matcher import PinnedBuffer
    if not ids:
        return
    size = max(m.indexes[i].header.body_size for i in ids.values())
    bufs = [PinnedBuffer(size) for _ in range(n_bufs)]
    pending = [None] * n_bufs

    def write(path, header, buf, n):
        tmp = f"{path}.tmp.{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(header)
            f.write(memoryview(buf.array)[:n])
        os.replace(tmp, path)

    with ThreadPoolExecutor(max_workers=n_bufs) as ex:
        for k, (name, idx) in enumerate(ids.items()):
            b = k % n_bufs
            if pending[b] is not None:
                pending[b].result()
            n = m.download_index_into(idx, bufs[b])
            pending[b] = ex.submit(write, os.path.join(outdir, f"{name}.cobs_classic"),
                                   m.indexes[idx].header.to_bytes(), bufs[b], n)
        for p in pending:
            if p is not None:
                p.result()


def expected_file_digests(m, records, res, mowner, names_in_rank_order, refs_by_rank):
    """sha256 of what match-db must leave on disk for this rank's indexes (decompressed text) and of the
    04_filter FASTA, formatted in-process from the device results."""
    from phylign_b200.cobs_text import format_cobs_text_fast, format_filter_fasta_fast
    out = {}
    for idx, ix in m.indexes.items():
        out[ix.batch] = hashlib.sha256(format_cobs_text_fast(records, res, ix, strip_prefix=True)).hexdigest()
    fa = None
    if mowner is not None:
        fa = hashlib.sha256(format_filter_fasta_fast([(h, s) for h, s in records], mowner.ptr, refs_by_rank)).hexdigest()
    return out, fa


def wait_gpus_released(n_gpus, timeout_s=20.0):
    """Set-up hygiene between one-shot runs: the previous holder of ~90 GB of HBM (this process' closed
    context, or the match-db run before) may still be tearing down; a CUDA context created meanwhile
    can take seconds instead of 0.3 s.  Wait until the driver reports the memory as free."""
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < timeout_s:
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=memory.used", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.split()
            used = [float(x) for x in out[:n_gpus]]
            if used and max(used) < 8000:       # MiB (this process keeps a small context of its own)
                break
        except Exception:
            break
        time.sleep(0.25)
    return time.perf_counter() - t0


def run_match_db(workdir, tag, fasta, n_gpus, bases, hbm_budget_gb=0, extra=()):
    """One files-in -> files-out run of the product CLI; returns the e2e_files record."""
    import gzip
    waited = wait_gpus_released(n_gpus)
    outdir = os.path.join(workdir, f"out_{tag}")
    tj = os.path.join(workdir, f"timing_{tag}.json")
    cmd = [sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", os.path.join(workdir, "cobs"),
           "--batches", os.path.join(workdir, "batches.txt"), "-q", fasta, "--qfile", "reads",
           "--match-dir", os.path.join(outdir, "03_match"), "--filter-out", os.path.join(outdir, "04_filter", "reads.fa"),
           "-t", str(THRESHOLD), "-n", str(TOP_N), "--timing-json", tj,
           "--benchmark-dir", os.path.join(outdir, "logs", "benchmarks", "run_cobs"),
           "--load-workers", str(min(16, host_cores()))]
    if n_gpus > 1:
        cmd += ["--gpus", str(n_gpus)]
    if hbm_budget_gb:
        cmd += ["--hbm-budget", str(int(hbm_budget_gb * 1e9))]
    cmd += list(extra)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE",
                                                            "MASTER_ADDR", "MASTER_PORT", "GROUP_RANK", "ROLE_RANK")}
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    t0 = time.perf_counter()
    try:        # a stuck run must not cost the bench line (the workers of --gpus N end with their parent)
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=ROOT, timeout=900)
    except subprocess.TimeoutExpired:
        shutil.rmtree(outdir, ignore_errors=True)
        return {"value": None, "error": "match-db did not finish within 900 s"}, None
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        shutil.rmtree(outdir, ignore_errors=True)
        return {"value": None, "error": r.stderr[-800:]}, None
    timing = json.load(open(tj))
    mdir = os.path.join(outdir, "03_match")
    files = sorted(os.listdir(mdir))
    got = {}
    for f in files:                 # decompressed content digests (the .gz bytes are not a parity target)
        with gzip.open(os.path.join(mdir, f), "rb") as g:
            got[f.split("____")[0]] = hashlib.sha256(g.read()).hexdigest()
    fa_path = os.path.join(outdir, "04_filter", "reads.fa")
    got_fa = hashlib.sha256(open(fa_path, "rb").read()).hexdigest()
    ph = timing["phases_s"]
    host_s = timing["total_s"] - ph.get("index_load_wait_s", 0.0) - ph.get("gpu_match_s", 0.0)
    rec = {"value": bases / wall, "unit": UNIT, "wall_s": round(wall, 3), "waited_for_free_hbm_s": round(waited, 2),
           "command": "python -m phylign_b200.cli match-db --cobs-dir DIR --batches FILE -q reads.fa --match-dir 03_match "
                      "--filter-out 04_filter/reads.fa -t 0.7 -n 100" + (f" --gpus {n_gpus}" if n_gpus > 1 else ""),
           "after_index_load": {"value": bases / max(1e-9, timing["total_s"] - ph.get("index_load_wait_s", 0.0) - ph.get("plan_s", 0.0)),
                                "unit": UNIT, "what": "the same run without index load and planning (queries -> files)"},
           "breakdown_s": dict(ph, process_start_and_exit=round(wall - timing["total_s"], 3), total_in_process=round(timing["total_s"], 3)),
           "host_s_outside_gpu_and_load": round(host_s, 3), "gpu_match_s": round(ph.get("gpu_match_s", 0.0), 3),
           "writer": timing["writer"], "direct_device_merge": timing.get("direct_device_merge"),
           "rounds": timing.get("rounds"), "overlap_rounds": timing.get("overlap_rounds"),
           "query_blocks": timing.get("query_blocks"), "filter_written_per_block": timing.get("filter_written_per_block"),
           "gpu_phase_ms_hash_gather_merge": timing.get("gpu_phase_ms_hash_gather_merge"),
           "inputs": {"index_files": len(files), "index_bytes": sum(os.path.getsize(os.path.join(workdir, "cobs", f))
                                                                     for f in os.listdir(os.path.join(workdir, "cobs"))),
                      "query_fasta_bytes": os.path.getsize(fasta), "index_format": ".cobs_classic (decompressed, /dev/shm)"},
           "outputs": {"match_files": len(files), "match_gz_bytes": sum(os.path.getsize(os.path.join(mdir, f)) for f in files),
                       "filter_fasta_bytes": os.path.getsize(fa_path),
                       "benchmark_logs": len(os.listdir(os.path.join(outdir, "logs", "benchmarks", "run_cobs")))}}
    shutil.rmtree(outdir, ignore_errors=True)
    return rec, (got, got_fa)


class _SkipFiles(Exception):
    pass


def run_ours(args, w, rank, world, local_rank):
    from phylign_b200 import _lib
    from phylign_b200.matcher import Matcher, PinnedBuffer, nccl_unique_id
    from phylign_b200.cobs_index import ref_of
    dist = Dist(rank, world, local_rank)
    m = Matcher(local_rank)
    if world > 1:
        m.nccl_init(dist.bcast(nccl_unique_id() if rank == 0 else None), rank, world)
        multi_gpu_options(m, args)
    specs = [_lib.SynthSpec(**spec_kwargs(b)) for b in w["batches"]]
    placement, imbalance = place(w, world, args.hbm_budget_gb * 10 ** 9)
    t_build = time.perf_counter()
    local = placement[rank]
    ids = {b["name"]: m.add_synth_index(b["name"], _lib.SynthSpec(**spec_kwargs(b)), b["signature_size"])
           for b in local}
    m.sync()
    t_build = time.perf_counter() - t_build
    all_names = [b["name"] for b in w["batches"]]
    m.set_ranks(all_names)
    L = w["read_len"]
    raw = m.synth_reads(specs, READS_SEED, 0, w["n_reads"], L, RANDOM_Q8, ERR_Q16)
    offs = np.arange(w["n_reads"] + 1, dtype=np.uint64) * L
    pin_raw, pin_offs = pinned_queries(raw, offs)    # the step's inputs live in pinned host memory
    local_row_bytes = sum((b["n_docs"] + 7) // 8 for b in local)
    local_alg_bytes = w["n_reads"] * w["kmers_per_read"] * local_row_bytes

    # ---- device-resident throughput (`value`): queries already in HBM
    m.set_queries_raw(raw, offs)
    main = timed_steps(dist, m, args.steps, args.warmup, rank, local_rank, sample_clocks=True)
    ms_per_step = main["ms_per_step"]
    value = w["bases"] / (ms_per_step * 1e-3)

    # ---- end to end through the public API: host buffers in, host results out, every step
    sharded = world > 1 and not args.merge_on_rank0
    e2e_s, h2d, d2h, e2e_parts, res, moffs, mcands = e2e_steps(dist, m, pin_raw, pin_offs, len(raw), offs.nbytes,
                                                               args.steps, args.warmup, sharded)
    h2d_all, d2h_all = dist.sum(h2d), dist.sum(d2h)             # whole job (all ranks)
    e2e_value = w["bases"] / e2e_s
    digest = result_digest(dist, m, res, moffs, mcands)
    n_units, n_hits, n_merged = int(dist.sum(len(res.units))), int(dist.sum(len(res.hits))), int(dist.sum(len(mcands)))
    want_files = (not args.no_e2e_files) and w["name"] == "reads1k"
    exp_files = exp_fa = None
    if want_files:      # what the files must hold, from these very device results
        names = [f"read_{i:07d}" for i in range(w["n_reads"])]
        records = [(names[i], raw[i * L:(i + 1) * L]) for i in range(w["n_reads"])]
        refs_by_rank = {br: [f"SYN{d:06d}" for d in range(b["n_docs"])]
                        for br, b in enumerate(sorted(w["batches"], key=lambda b: b["name"]))}
        fo, fc = full_merged(dist, m, moffs, mcands)
        mstruct = keep = None
        if rank == 0:
            import ctypes as C
            mstruct, keep = merged_struct(fo, fc)
            mstruct = type("P", (), {"ptr": C.pointer(mstruct)})()
        exp_files, exp_fa = expected_file_digests(m, records, res, mstruct, all_names, refs_by_rank)
        del records, keep, fo, fc
    del res, moffs, mcands

    # ---- the same gather with the exact threshold pruning switched off (explains the roofline)
    m.set_option("prune", 0)
    unpruned_ms = []
    for i in range(2 + 3):
        m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
        if i >= 2:
            unpruned_ms.append(m.phase_ms()[1])
    res0 = m.fetch()
    mo0, mc0 = m.merged()
    digest_unpruned = result_digest(dist, m, res0, mo0, mc0)
    del res0, mo0, mc0
    m.set_option("prune", 1)
    dist.barrier()
    if rank == 0 and digest != digest_unpruned:
        raise SystemExit(f"result_digest differs with pruning off: {digest} vs {digest_unpruned}")

    # ---- secondary: the short-read class (150 bp, the length of the reference's own test reads)
    secondary = {}
    w150 = with_read_len(w, SHORT_READ_LEN)
    raw150 = m.synth_reads(specs, READS_SEED, 0, w150["n_reads"], SHORT_READ_LEN, RANDOM_Q8, ERR_Q16)
    raw4 = None
    if want_files and world == 1:      # 4 x the reads for the multi-block file run
        raw4 = m.synth_reads(specs, READS_SEED, 0, 4 * w["n_reads"], L, RANDOM_Q8, ERR_Q16)
    offs150 = np.arange(w150["n_reads"] + 1, dtype=np.uint64) * SHORT_READ_LEN
    m.set_queries_raw(raw150, offs150)
    s150 = timed_steps(dist, m, max(3, min(args.steps, 10)), 3, rank, local_rank)
    m.set_option("prune", 0)
    m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
    m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
    s150_unpruned_ms = m.phase_ms()[1]
    m.set_option("prune", 1)
    peak, peak_src = measured_peak()
    alg150 = w150["n_reads"] * w150["kmers_per_read"] * local_row_bytes
    secondary["reads150"] = {
        "workload": config_dict(w150)["workload"], "value": w150["bases"] / (s150["ms_per_step"] * 1e-3), "unit": UNIT,
        "ms_per_step": s150["ms_per_step"], "phase_ms_hash_gather_merge": s150["phases"],
        "roofline": {"kernel": "gather_count_ring_kernel<32,8,2,4>", "bytes_per_launch": s150["gathered"],
                     "achieved": s150["gathered"] / (s150["gather_ms"] * 1e-3) / 1e9,
                     "frac": s150["gathered"] / (s150["gather_ms"] * 1e-3) / 1e9 / peak,
                     "unpruned": {"ms": float(s150_unpruned_ms), "frac": alg150 / (s150_unpruned_ms * 1e-3) / 1e9 / peak}}}
    dist.barrier()

    # ---- CPU baseline beside it (N=1 only): the oracle on the host cores, bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            native = oracle.build_native()
            cores = host_cores()
            n_sample = min(w["n_reads"], args.cpu_sample_reads)
            w0 = w["batches"][0]
            oidx = oracle.OracleIndex.new(w0["n_docs"], w0["signature_size"])
            m.download_index_into(ids[w0["name"]], oidx.body.reshape(-1))
            reads = [raw[r * L:(r + 1) * L] for r in range(n_sample)]
            lay = time_oracle_layouts(oidx, reads, cores)
            best = min(lay, key=lay.get)
            v = n_sample * L / (lay[best] * w["n_indexes"])
            cpu_baseline = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{n_sample} reads against 1 of the {w['n_indexes']} indexes ({lay[best]:.2f} s), extrapolated "
                          f"x{w['n_indexes']}; oracle port of cobs 0.2.1 classic query (NOT the cobs binary), "
                          f"{'-O3 -march=native' if native else '-O3 -march=x86-64-v2'}, kernel {best[0]}, layout '{best[1]}'",
                "variants_s": {f"{k[0]} | {k[1]}": round(t, 3) for k, t in lay.items()},
                "cobs_on_path": find_cobs()}
            del oidx
        except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
            cpu_baseline = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": f"failed: {e}"}

    # ---- files in -> files out through the product CLI
    e2e_files = None
    workdir = None
    if want_files:      # room for the index files?  (a bench line without e2e_files beats no line at all)
        if rank == 0:
            workdir = shm_dir("phylign_bench_")
            need = sum(b["signature_size"] * ((b["n_docs"] + 7) // 8) for b in w["batches"]) + 8 * len(raw) + (4 << 30)
            free = shutil.disk_usage(workdir).free
            if free < need:
                e2e_files = {"value": None, "error": f"{workdir}: {free / 1e9:.0f} GB free, {need / 1e9:.0f} GB needed for "
                                                     f"the index files: files-in -> files-out run skipped"}
                shutil.rmtree(workdir, ignore_errors=True)
                workdir = None
        workdir = dist.bcast(workdir)
        want_files = workdir is not None
    if want_files:
        try:
            os.makedirs(os.path.join(workdir, "cobs"), exist_ok=True)
            t_w = time.perf_counter()
            failed = 0
            try:
                write_index_files(m, ids, os.path.join(workdir, "cobs"))
                if rank == 0:
                    with open(os.path.join(workdir, "batches.txt"), "w") as f:
                        f.write("\n".join(all_names) + "\n")
                    write_fasta(os.path.join(workdir, "reads.fa"), raw, w["n_reads"], L)
                    write_fasta(os.path.join(workdir, "reads150.fa"), raw150, w150["n_reads"], SHORT_READ_LEN)
            except OSError as e:
                failed = 1
                e2e_files = {"value": None, "error": f"writing the bench inputs failed: {e}"}
            t_w = time.perf_counter() - t_w
            if dist.max(failed):
                e2e_files = e2e_files or {"value": None, "error": "writing the bench inputs failed on another rank"}
                raise _SkipFiles()
            if world > 1:
                m.nccl_finalize()           # collective, orderly end of the communicator
            m.close()                       # the CLI needs the HBM
            m = None
            exp_all = dist.gather(exp_files)
            dist.barrier()
            if rank == 0:
                e2e_files, got = run_match_db(workdir, "1k", os.path.join(workdir, "reads.fa"), world, w["bases"])
                if got is not None:       # a one-shot wall clock on a shared host is noisy: best of two runs
                    again, got2 = run_match_db(workdir, "1k_b", os.path.join(workdir, "reads.fa"), world, w["bases"])
                    runs = [e2e_files["wall_s"]] + ([again["wall_s"]] if got2 is not None else [])
                    if got2 is not None and got2 == got and again["wall_s"] < e2e_files["wall_s"]:
                        e2e_files = again
                    e2e_files["wall_s_runs"] = runs
                if got is not None:
                    want = {}
                    for p in exp_all:
                        want.update(p)
                    e2e_files["files_equal_device_results"] = bool(got[0] == want and got[1] == exp_fa)
                    e2e_files["setup_write_index_files_s"] = round(t_w, 1)
                    if not e2e_files["files_equal_device_results"]:
                        bad = [b for b in want if got[0].get(b) != want[b]]
                        raise SystemExit(f"match-db files differ from the in-process device results: "
                                         f"{len(bad)} match files, 04_filter equal: {got[1] == exp_fa}")
                    f150, _ = run_match_db(workdir, "150", os.path.join(workdir, "reads150.fa"), world, w150["bases"])
                    secondary["reads150"]["e2e_files"] = f150
                    if world == 1 and raw4 is not None:
                        # steady state: 4 query blocks of 100k reads in one run -- the match files of block i are
                        # formatted, gzipped and appended, and its slice of 04_filter written, while the GPU matches
                        # block i+1
                        write_fasta(os.path.join(workdir, "reads4.fa"), raw4, 4 * w["n_reads"], L)
                        f4, _ = run_match_db(workdir, "4blocks", os.path.join(workdir, "reads4.fa"), world, 4 * w["bases"],
                                             extra=["--query-block-bases", str(w["bases"])])
                        if f4.get("value"):
                            b4 = f4["breakdown_s"]
                            after = b4["total_in_process"] - b4.get("index_load_wait_s", 0) - b4.get("ctx_create_s", 0) - b4.get("plan_s", 0)
                            f4["per_block_s_after_load"] = round(after / 4, 3)
                            f4["per_block_gpu_s"] = round(b4.get("gpu_match_s", 0) / 4, 3)
                            f4["bases_per_s_after_load"] = 4 * w["bases"] / after
                        secondary["files_4_query_blocks"] = f4
                    # HBM-overflow streaming: the same run with the per-GPU budget capped so that the
                    # batches need several resident rounds; round r+1 loads while round r is matched
                    per_gpu = sum(b["signature_size"] * 512 for b in w["batches"]) / world
                    cap_gb = max(8.0, 0.55 * per_gpu / 1e9 + 4.0)
                    fcap, gotc = run_match_db(workdir, "capped", os.path.join(workdir, "reads.fa"), world, w["bases"],
                                              hbm_budget_gb=cap_gb)
                    if gotc is not None:
                        fcap["files_equal_device_results"] = bool(gotc[0] == want and gotc[1] == exp_fa)
                        fcap["hbm_budget_gb_per_gpu"] = round(cap_gb, 1)
                        fcap["wall_vs_fully_resident"] = round(fcap["wall_s"] / e2e_files["wall_s"], 3)
                        if not fcap["files_equal_device_results"]:
                            raise SystemExit("capped-budget match-db run produced different files")
                    secondary["streaming_capped_hbm"] = fcap
            dist.barrier()
        except _SkipFiles:
            pass
        finally:
            dist.barrier()
            if rank == 0:
                shutil.rmtree(workdir, ignore_errors=True)

    # ---- at 8 GPUs: BASELINE configs[3], the full 661k-shaped database
    if (world == 8 or (args.force_db661k and world > 1)) and w["name"] == "reads1k" and not args.no_db661k:
        if m is not None:
            m.nccl_finalize()
            m.close()
            m = None
        sec = secondary_db661k(args, dist, rank, world, local_rank)
        if rank == 0:
            secondary.update(sec)
    if m is not None and world > 1:
        m.nccl_finalize()
    if rank != 0:
        if m is not None:
            m.close()
        return
    kernel = ("gather_count_ring_kernel<32,10,3,4>" if w["name"] == "reads1k" else
              "gather_count_ring_kernel<LPR,10,3,4> (one launch per row-width class)")
    traffic, traffic_src = None, None
    try:   # DRAM bytes per launch from the committed ncu capture of this exact workload AND kernel source, else null
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = f"{kernel}|{w['name']}|{w['n_reads']}|{w['read_len']}|{w['n_indexes']}|{w['n_docs']}|{world}|prune1"
        ent = tj.get(key)
        if ent and ent.get("kernel_source_sha256") == kernel_source_digest():
            traffic, traffic_src = ent["traffic_bytes"], ent.get("source")
    except Exception:
        pass
    g_ms = main["gather_ms"]
    gathered = main["gathered"]
    achieved = local_alg_bytes / (g_ms * 1e-3) / 1e9
    un_ms = float(np.mean(unpruned_ms))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": config_dict(w),
            "run": {"sharding": f"indexes placed LPT-by-row-bytes on {world} GPU(s) (imbalance {imbalance:.3f}), "
                                f"queries replicated" + ("" if world == 1 else
                                                         (", merged lists on rank 0" if args.merge_on_rank0 else
                                                          f"; each rank uploads 1/{world} of the bases (NVLink all-gather) and "
                                                          f"finalises + downloads 1/{world} of the queries' merged lists")),
                    "index_build_s": round(t_build, 2),
                    "kmer_docs_per_s": w["kmer_docs"] / (ms_per_step * 1e-3),
                    "phase_ms_hash_gather_merge": main["phases"],
                    "pruning": "exact threshold pruning on (output-identical: result_digest equals the unpruned run's); "
                               "roofline.unpruned = the same step with every row read",
                    "n_units": n_units, "n_hits": n_hits, "n_merged": n_merged},
            "result_digest": digest,
            "result_digest_checks": {"equals_unpruned_run": digest == digest_unpruned,
                                     "what": "sha256 over per-batch (query, n_pass, n_kept, hits) in batch order + merged "
                                             "lists; independent of N and placement"},
            # units one launch processes = (k-mer, index) pairs whose row was gathered; the exact
            # threshold pruning ends a (query,index) unit once no document can reach -t any more, so
            # that is fewer pairs than K_q x indexes.  achieved = gathered row bytes / time.
            "roofline": {"bound": "hbm", "achieved": gathered / (g_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": gathered / (g_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": kernel,
                         "bytes_per_launch": int(gathered),
                         "all_pairs": {"bytes_per_launch": int(local_alg_bytes), "achieved": achieved,
                                       "frac": achieved / peak,
                                       "what": "SURVEY 8(d) bytes of ALL (k-mer,index) pairs of the step / the same "
                                               "time: > peak because pairs that cannot change the output are "
                                               "never gathered (results are bit-identical, tests/ run with "
                                               "pruning on)"},
                         "unpruned": {"ms": un_ms, "achieved": local_alg_bytes / (un_ms * 1e-3) / 1e9,
                                      "frac": local_alg_bytes / (un_ms * 1e-3) / 1e9 / peak,
                                      "what": "same launch with pruning off: every pair gathered"},
                         "note": f"row bytes gathered by rank 0's launch / mean CUDA-event duration of the "
                                 f"gather phase ({g_ms:.2f} ms, one launch per step); peak = {peak_src}"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                    "ms_per_step": e2e_s * 1e3, "breakdown_ms_set_run_fetch_merged": e2e_parts,
                    "bytes_are": "summed over all ranks (each rank copies its share of the queries / results)"},
            "e2e_files": e2e_files,
            "secondary": secondary,
            "gpu_launches": main["launches"], "clocks": main["clocks"]}
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    exp = expected_digest(w, args)
    if exp is not None:
        line["result_digest_checks"]["equals_committed_expectation"] = digest == exp
        if digest != exp:
            print(json.dumps(line))
            raise SystemExit(f"result_digest {digest} != committed expectation {exp} (profiles/expected_digests.json)")
    if m is not None:
        m.close()
    print(json.dumps(line))


def kernel_source_digest():
    h = hashlib.sha256()
    for f in ("gather_count.cu", "phy_internal.cuh"):
        h.update(open(os.path.join(ROOT, "phylign_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def expected_digest(w, args):
    """The digest this workload produced when it was last verified (any N): lets every run, at any
    number of GPUs, assert that it reproduces the same results."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "expected_digests.json")))
        b0 = w["batches"][0]
        key = f"{w['name']}|{w['n_reads']}|{w['read_len']}|{w['n_indexes']}|{w['n_docs']}|{b0['genome_len']}|{getattr(args, 'db_scale', 1.0)}"
        return d.get(key)
    except Exception:
        return None


def secondary_db661k(args, dist, rank, world, local_rank):
    """BASELINE configs[3] on 8 GPUs: the 305 batches of the 661k database (1.06 TB) resident across the
    GPUs' HBM, the same 100k x 1 kbp reads.  A few steps; digest asserted pruning on == off."""
    from phylign_b200 import _lib
    from phylign_b200.matcher import Matcher, nccl_unique_id
    a2 = argparse.Namespace(**vars(args))
    a2.workload, a2.db_scale = "db661k", args.db_scale
    w = workload(a2)
    m = Matcher(local_rank)
    m.nccl_init(dist.bcast(nccl_unique_id() if rank == 0 else None), rank, world)
    multi_gpu_options(m, args)
    placement, imbalance = place(w, world, args.hbm_budget_gb * 10 ** 9)
    t0 = time.perf_counter()
    for b in placement[rank]:
        m.add_synth_index(b["name"], _lib.SynthSpec(**spec_kwargs(b)), b["signature_size"])
    m.sync()
    t_build = dist.max(time.perf_counter() - t0)
    m.set_ranks([b["name"] for b in w["batches"]])
    specs = [_lib.SynthSpec(**spec_kwargs(b)) for b in w["batches"]]
    L = w["read_len"]
    raw = m.synth_reads(specs, READS_SEED, 0, w["n_reads"], L, RANDOM_Q8, ERR_Q16)
    offs = np.arange(w["n_reads"] + 1, dtype=np.uint64) * L
    m.set_queries_raw(raw, offs)
    steps = 5
    t = timed_steps(dist, m, steps, 3, rank, local_rank)
    pin_raw, pin_offs = pinned_queries(raw, offs)
    e2e_s, h2d, d2h, parts, res, moffs, mcands = e2e_steps(dist, m, pin_raw, pin_offs, len(raw), offs.nbytes, steps, 2,
                                                           not args.merge_on_rank0)
    h2d, d2h = dist.sum(h2d), dist.sum(d2h)
    dg = result_digest(dist, m, res, moffs, mcands)
    del res, moffs, mcands
    m.set_option("prune", 0)
    m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
    m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
    un_ms = dist.max(m.phase_ms()[1])
    res0 = m.fetch()
    mo0, mc0 = m.merged()
    dg0 = result_digest(dist, m, res0, mo0, mc0)
    del res0, mo0, mc0
    m.set_option("prune", 1)
    local_row_bytes = sum((b["n_docs"] + 7) // 8 for b in placement[rank])
    m.nccl_finalize()
    m.close()
    dist.barrier()
    if rank != 0:
        return {}
    if dg != dg0:
        raise SystemExit(f"db661k: result_digest differs with pruning off: {dg} vs {dg0}")
    exp = expected_digest(w, a2)
    if exp is not None and dg != exp:
        raise SystemExit(f"db661k: result_digest {dg} != committed expectation {exp} (profiles/expected_digests.json)")
    peak, _ = measured_peak()
    alg_local = w["n_reads"] * w["kmers_per_read"] * local_row_bytes
    return {"db661k": {"workload": config_dict(w)["workload"], "value": w["bases"] / (t["ms_per_step"] * 1e-3),
                       "unit": UNIT, "ms_per_step": t["ms_per_step"], "steps": steps,
                       "e2e": {"value": w["bases"] / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                               "d2h_bytes_per_step": int(d2h)},
                       "index_build_s": round(t_build, 2), "placement_imbalance": round(imbalance, 4),
                       "phase_ms_hash_gather_merge": t["phases"],
                       "algorithmic_bytes_per_step": w["alg_bytes"],
                       "aggregate_algorithmic_TBps": w["alg_bytes"] / (t["ms_per_step"] * 1e-3) / 1e12,
                       "roofline_rank0": {"bytes_gathered": t["gathered"],
                                          "frac": t["gathered"] / (t["gather_ms"] * 1e-3) / 1e9 / peak,
                                          "unpruned_frac": alg_local / (un_ms * 1e-3) / 1e9 / peak},
                       "result_digest": dg, "equals_unpruned_run": True,
                       "equals_committed_expectation": (dg == exp) if exp is not None else None}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-setup"])
    ap.add_argument("--reads", type=int, default=100_000)
    ap.add_argument("--read-len", type=int, default=1000)
    ap.add_argument("--indexes", type=int, default=64)
    ap.add_argument("--docs", type=int, default=4000)
    ap.add_argument("--genome-len", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-reads", type=int, default=20_000)
    ap.add_argument("--workload", default="reads1k", choices=["reads1k", "db661k"])
    ap.add_argument("--db-scale", type=float, default=1.0, help="db661k: scale every signature_size")
    ap.add_argument("--hbm-budget-gb", type=float, default=170.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-files", action="store_true", help="skip the files-in -> files-out run of match-db")
    ap.add_argument("--no-db661k", action="store_true", help="at --gpus 8: skip secondary.db661k")
    ap.add_argument("--force-db661k", action="store_true", help=argparse.SUPPRESS)   # secondary.db661k at any N > 1 (tests)
    ap.add_argument("--merge-on-rank0", action="store_true",
                    help="N > 1: gather the merged lists on rank 0 and replicate the query upload (round-1 behaviour)")
    ap.add_argument("--workdir", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    args.gpus = max(args.gpus, world)
    w = workload(args)
    if args.impl == "reference-setup":
        sys.exit(cmd_reference_setup(args, w))
    if args.impl == "reference":
        run_reference(args, w, rank, world)
    else:
        run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
