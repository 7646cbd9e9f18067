"""Command-line drop-ins for the reference's match-stage commands.

    python -m phylign_b200.cli cobs query [--load-complete] -t THR -T N -i INDEX
                                          [--index-sizes BYTES] -f QUERY.fa
        same flags and stdout protocol as the `cobs query` call of
        /root/reference/scripts/run_cobs_streaming.sh:24-29 and Snakefile:419-424,476-481
        (-T is accepted and ignored: the GPU replaces the thread pool)
    python -m phylign_b200.cli run-cobs-streaming THR THREADS INDEX.xz SIZE QUERY.fa
        the 5 positionals of scripts/run_cobs_streaming.sh:13-22
    python -m phylign_b200.cli postprocess -n N           (stdin -> stdout)
        scripts/postprocess_cobs.py:42-58
    python -m phylign_b200.cli filter -n N -q QUERY.fa MATCH.gz [MATCH.gz ...]
        scripts/filter_queries.py:209-238 (merge on the GPU, FASTA on stdout, log on stderr)
    python -m phylign_b200.cli match-db --cobs-dir DIR --batches FILE -q QUERY.fa
                                        --match-dir intermediate/03_match --filter-out OUT.fa
        all batches in one resident context: writes every {batch}____{qfile}.gz and the
        04_filter FASTA (replaces 305 `decompress_and_run_cobs` jobs + `translate_matches`);
        --shard I/N (one process per GPU), --round-bytes (stream batches that do not fit),
        --resume, --bucket-dir (reference -> queries tables for stage 05)
    python -m phylign_b200.cli serve --socket SOCK [--preload INDEX ...]
        resident server: indexes stay in HBM, `cobs query --server SOCK` / $PHYLIGN_SERVER
    python -m phylign_b200.cli fix-query INPUT.f[aq] ...
        seqtk seq -A -U -C | awk non-ACGT->A of Snakefile:326-332, inputs concatenated

Any failure exits non-zero and leaves nothing partial at the output paths (the rules run
under `set -euo pipefail`, Snakefile:142).
"""
from __future__ import annotations

import argparse
import gzip
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import fasta
from .cobs_index import ref_of as _ref_of
from .cobs_text import format_filter_fasta_fast


def _die(msg, code=1):
    print(f"phylign_b200: error: {msg}", file=sys.stderr)
    sys.exit(code)


def _postprocess_stream(fin, fout, keep: int):
    """Drop-in for `postprocess_cobs.py -n keep` (scripts/postprocess_cobs.py:21-38) on a text stream
    (for real `cobs query` output; our own `cobs query --top-n` fuses this on the GPU).  Per query
    block: names become "_" + text after the first "_", the first `keep` hit lines are kept plus
    every later line whose count equals that of line number `keep` (cobs sorts by score, so those
    are the ties).  keep <= 0 keeps the lines with count 0 only, as the script does."""
    def flush(block):
        if not block:
            return
        lines = ["_" + x.partition("_")[2] for x in block]
        if 0 < keep < len(lines):
            cut = int(lines[keep - 1].split("\t")[-1])
            lines = lines[:keep] + [y for y in lines[keep:] if int(y.split("\t")[-1]) == cut]
        elif keep <= 0:
            lines = [y for y in lines if int(y.split("\t")[-1]) == 0]
        fout.writelines(lines)

    block = []
    for x in fin:
        if x[0] == "*":
            flush(block)
            block = []
            fout.write(x)
        else:
            block.append(x)
    flush(block)


# ------------------------------------------------------------------------------------ cobs query
def cmd_cobs_query(a):
    server = getattr(a, "server", None) or os.environ.get("PHYLIGN_SERVER")
    if server:      # resident indexes: ship the request to `phylign_b200.cli serve`
        from .server import request
        head, payload = request(server, {"cmd": "query", "index": os.path.abspath(a.i), "query": os.path.abspath(a.f),
                                         "threshold": a.t, "top_n": a.top_n, "floor": a.floor,
                                         "index_sizes": a.index_sizes})
        if not head.get("ok"):
            _die(head.get("error", "server error"))
        sys.stdout.buffer.write(payload)
        sys.stdout.buffer.flush()
        return
    from .matcher import Matcher
    from .cobs_text import format_cobs_text_arrays
    qf = fasta.QueryFile(a.f)                                # flat arrays (native reader)
    with Matcher(a.device) as m:
        if getattr(a, "sanitize_queries", False):
            m.set_option("sanitize_queries", 1)
        m.set_option("pinned_results", 0)                    # one fetch per block: plain host memory is cheaper
        idx = m.load_index(a.i, batch="index")
        hdr = m.indexes[idx].header
        if a.index_sizes is not None and a.index_sizes != hdr.header_size + hdr.body_size:
            _die(f"--index-sizes {a.index_sizes} != header {hdr.header_size} + body {hdr.body_size}")
        for q0, q1 in qf.block_ranges(a.query_block_bases):
            m.set_queries_raw(qf.seqs, qf.soffs[q0:q1 + 1])
            res = m.match(a.t, top_n=a.top_n, floor_mode=a.floor)
            sys.stdout.buffer.write(format_cobs_text_arrays(qf.headers, qf.hoffs[q0:q1 + 1], res, m.indexes[idx],
                                                            strip_prefix=a.top_n > 0))
        sys.stdout.buffer.flush()
        m.release_at_exit()



# ------------------------------------------------------------------------------------ filter
def _fixed_width(buf: np.ndarray, off: np.ndarray, ln: np.ndarray, width: int = 0) -> np.ndarray:
    """Substrings buf[off[i] : off[i]+ln[i]] as one NUL-padded fixed-width bytes array (dtype S<width>),
    gathered with numpy -- no per-string Python objects."""
    if len(off) == 0:
        return np.zeros(0, dtype=f"S{max(1, width)}")
    width = max(width, int(ln.max()), 1)
    cols = np.arange(width, dtype=np.int64)
    idx = np.minimum(off[:, None] + cols[None, :], len(buf) - 1)
    mat = np.where(cols[None, :] < ln[:, None], buf[idx], 0).astype(np.uint8)
    return np.ascontiguousarray(mat).view(f"S{width}").ravel()


def parse_match_file_native(path):
    """A match file parsed with the rules of filter_queries.py:27-66 by the library's C++ parser, as arrays:
    (qnames S-array per block, first_hit uint64[n_blocks+1], ref_ids int64[n_hits], refs_sorted [str],
    kmers uint32[n_hits]); ref_ids index refs_sorted (byte order = Python str order for ASCII)."""
    import ctypes as C
    from . import _lib
    L = _lib.load()
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        text = f.read()
    mp = C.POINTER(_lib.MatchText)()
    rc = L.phy_parse_match_text(text, len(text), C.byref(mp))
    if rc != 0:
        raise ValueError(f"{path}: {L.phy_last_error(None).decode()}")
    try:
        m = mp.contents
        nb, nh = int(m.n_blocks), int(m.n_hits)
        arr = lambda p, n, dt: np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)
        q_off, q_len = arr(m.q_off, nb, np.int64), arr(m.q_len, nb, np.int64)
        first_hit = arr(m.first_hit, nb + 1, np.uint64)
        ref_off, ref_len = arr(m.ref_off, nh, np.int64), arr(m.ref_len, nh, np.int64)
        kmers = arr(m.kmers, nh, np.uint32)
    finally:
        L.phy_match_text_free(mp)
    buf = np.frombuffer(text, dtype=np.uint8)
    qnames = _fixed_width(buf, q_off, q_len)
    if nh:
        uniq, inv = np.unique(_fixed_width(buf, ref_off, ref_len), return_inverse=True)
        refs_sorted = [u.decode() for u in uniq.tolist()]
        ref_ids = inv.astype(np.int64)
    else:
        refs_sorted, ref_ids = [], np.zeros(0, np.int64)
    return qnames, first_hit, ref_ids, refs_sorted, kmers


class _QueryNameIndex:
    """Query name -> position in the query file, for whole arrays of names at once.  The usual match
    file lists the queries of the query file in order: that case is one array comparison."""

    def __init__(self, qid: dict):
        names = [None] * len(qid)
        for n, i in qid.items():
            names[i] = n.encode()
        self.width = max([len(n) for n in names] + [1])
        self.names = np.array(names, dtype=f"S{self.width}") if names else np.zeros(0, dtype="S1")
        self.order = np.argsort(self.names, kind="stable")
        self.sorted = self.names[self.order]

    def lookup(self, qnames: np.ndarray, what: str) -> np.ndarray:
        if len(qnames) == len(self.names) and qnames.dtype.itemsize <= self.width and \
                bool((qnames.astype(self.names.dtype) == self.names).all()):
            return np.arange(len(self.names), dtype=np.int64)
        q = qnames.astype(f"S{max(self.width, qnames.dtype.itemsize)}")
        srt = self.sorted.astype(q.dtype)
        pos = np.minimum(np.searchsorted(srt, q), max(len(srt) - 1, 0))
        ok = (srt[pos] == q) if len(srt) else np.zeros(len(q), bool)
        if not bool(ok.all()):
            bad = q[~ok][0].decode(errors="replace")
            raise KeyError(f"query {bad!r} of {what} is not in the query file")
        return self.order[pos].astype(np.int64)


def _load_filter_queries(query_fn):
    """(ordered {qname: seq}, {qname: position}) the way filter_queries.py:163-176 builds its dict."""
    queries = {}
    for qname, seq in fasta.read_fastx(query_fn):
        queries[qname] = seq                      # duplicate names overwrite, first position kept
    return queries, {q: i for i, q in enumerate(queries)}


def _parsed_pieces(match_fns, qid, brank, log):
    """Candidates of already written match files: [(qid array, CAND array)], {batch_rank: refs}.
    Files are parsed natively (phy_parse_match_text); several files of one batch are allowed."""
    from .matcher import CAND_DT
    by_batch = {}
    with ThreadPoolExecutor(max_workers=min(16, max(1, len(match_fns)))) as ex:   # gunzip + C++ parser release the GIL
        parsed_files = list(ex.map(parse_match_file_native, match_fns))
    for fn, parsed_one in zip(match_fns, parsed_files):
        batch = os.path.basename(fn).split("____")[0]
        print(f"Translating matches {fn}", file=log)
        by_batch.setdefault(batch, []).append(parsed_one)
    pieces, refs_by_rank = [], {}
    qindex = _QueryNameIndex(qid)
    for batch, parsed in by_batch.items():
        br = brank[batch]
        refs = sorted({r for p in parsed for r in p[3]})          # accessions of the batch, str order
        rr = {r: i for i, r in enumerate(refs)}
        refs_by_rank[br] = refs
        for qnames, first_hit, ref_ids, refs_sorted, kmers in parsed:
            q_of_block = qindex.lookup(qnames, f"batch {batch}")
            remap = np.array([rr[r] for r in refs_sorted], dtype=np.int64)   # file-local id -> batch rank
            rank = remap[ref_ids] if len(ref_ids) else np.zeros(0, np.int64)
            c = np.zeros(len(kmers), dtype=CAND_DT)
            c["score"], c["batch_rank"], c["doc"], c["ref_rank"] = kmers, br, rank, rank
            counts = np.diff(first_hit.astype(np.int64))
            pieces.append((np.repeat(q_of_block, counts), c))
    return pieces, refs_by_rank


def _merge_pieces(m, nq: int, pieces, keep: int):
    """Global top-N + ties over candidate pieces (filter_queries.py:123-150) on the GPU; the result is
    the Matcher's current merged list."""
    from .matcher import CAND_DT
    qs = np.concatenate([p[0] for p in pieces]) if pieces else np.zeros(0, np.int64)
    cs = np.concatenate([p[1] for p in pieces]) if pieces else np.zeros(0, CAND_DT)
    order = np.argsort(qs, kind="stable")
    offs = np.zeros(nq + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(np.bincount(qs, minlength=nq), dtype=np.uint64)
    return m.merge_host(offs, cs[order], keep)


def _final_merge(m, queries, pieces, refs_by_rank, keep: int) -> str:
    _merge_pieces(m, len(queries), pieces, keep)
    return format_filter_fasta_fast(list(queries.items()), m._merged_owner.ptr, refs_by_rank).decode()


def merge_match_files(m, query_fn, match_fns, keep: int, log=sys.stderr) -> str:
    """filter_queries.py process_files on the GPU (phy_merge_host)."""
    queries, qid = _load_filter_queries(query_fn)
    batch_names = sorted({os.path.basename(fn).split("____")[0] for fn in match_fns})
    brank = {b: i for i, b in enumerate(batch_names)}
    pieces, refs_by_rank = _parsed_pieces(match_fns, qid, brank, log)
    return _final_merge(m, queries, pieces, refs_by_rank, keep)


def cmd_filter(a):
    from .matcher import Matcher
    with Matcher(a.device) as m:
        out = merge_match_files(m, a.query_fn, a.match_fn, a.keep)
    sys.stdout.write(out)
    sys.stdout.flush()


def cmd_fix_query(a):
    """rules fix_query + concatenate_queries (Snakefile:314-352): FASTA/FASTQ in, one-line FASTA out,
    names without comments, bases upper-case ACGT (everything else -> A)."""
    if a.device is None:
        sys.stdout.write("".join(fasta.fix_query_file(p) for p in a.inputs))
        return
    from .matcher import Matcher
    recs = [r for p in a.inputs for r in fasta.read_fastx(p)]
    cat = np.frombuffer(bytearray("".join(s for _, s in recs).encode()), dtype=np.uint8)
    with Matcher(a.device) as m:
        m.fix_bases(cat)
    out, pos = [], 0
    fixed = cat.tobytes().decode()
    for name, s in recs:
        out.append(f">{name}\n{fixed[pos:pos + len(s)]}\n")
        pos += len(s)
    sys.stdout.write("".join(out))


# ------------------------------------------------------------------------------------ whole database
def _atomic_write(path, data: bytes, gz: bool):
    tmp = f"{path}.tmp.{os.getpid()}"
    try:
        if gz:
            with gzip.GzipFile(tmp, "wb", compresslevel=1, mtime=0) as f:
                f.write(data)
        else:
            with open(tmp, "wb") as f:
                f.write(data)
        os.replace(tmp, path)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)


def _spawn_match_db_workers(a):
    """`match-db --gpus N`: one worker process per GPU (device = rank), NCCL merge on rank 0.
    The parent supervises: the first worker that exits non-zero (bad index, PHY_ERR_NOMEM, a query
    outside ACGT ...) takes the others down, which would otherwise block forever inside a
    collective waiting for it; the job then exits non-zero (`set -euo pipefail` callers)."""
    import shutil
    import signal
    import subprocess
    import tempfile
    import threading
    import time
    if a.filter_out:                           # parts of an earlier, interrupted run must not be joined
        import glob
        for p in glob.glob(glob.escape(a.filter_out) + ".part.*"):
            os.unlink(p)
    tmpdir = tempfile.mkdtemp(prefix="phylign_nccl_")
    id_file = os.path.join(tmpdir, "id")
    argv = [x for x in sys.argv[1:]]
    procs = []

    def _stop(signum, _frame):                 # a cancelled job (snakemake, scancel, ^C) takes its workers along
        raise SystemExit(128 + signum)
    old_handlers = {}
    if threading.current_thread() is threading.main_thread():
        for sg in (signal.SIGTERM, signal.SIGINT, signal.SIGHUP):
            old_handlers[sg] = signal.signal(sg, _stop)
    try:
        for r in range(a.gpus):
            env = dict(os.environ, PHYLIGN_RANK=str(r), PHYLIGN_WORLD=str(a.gpus), PHYLIGN_NCCL_ID_FILE=id_file)
            procs.append(subprocess.Popen([sys.executable, "-m", "phylign_b200.cli"] + argv, env=env))
        failed = None
        while failed is None and any(p.poll() is None for p in procs):
            for r, p in enumerate(procs):
                if p.poll() not in (None, 0):
                    failed = r
                    break
            else:
                time.sleep(0.05)
        if failed is None:
            failed = next((r for r, p in enumerate(procs) if p.returncode != 0), None)
        if failed is not None:
            for p in procs:
                if p.poll() is None:
                    p.terminate()
            deadline = time.time() + 10
            for p in procs:
                try:
                    p.wait(timeout=max(0.1, deadline - time.time()))
                except subprocess.TimeoutExpired:
                    p.kill()
                    p.wait()
            _die(f"match-db worker {failed} failed (exit code {procs[failed].returncode}); "
                 f"the other workers were stopped")
        if a.filter_out:                       # query-sharded merge: every worker left its slices as parts
            import glob
            parts = [p for p in glob.glob(glob.escape(a.filter_out) + ".part.*") if ".tmp." not in p]
            if parts:
                _join_parts(a.filter_out, parts)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        for p in procs:
            p.wait()
        shutil.rmtree(tmpdir, ignore_errors=True)
        for sg, h in old_handlers.items():
            signal.signal(sg, h)


def _die_with_parent():
    """Worker side of `match-db --gpus N`: SIGTERM when the supervising parent goes away without being
    able to stop its workers (SIGKILL, OOM killer) -- an orphan would sit in an NCCL collective holding
    its GPU.  Linux prctl(PR_SET_PDEATHSIG); a no-op where prctl is missing."""
    import ctypes
    import signal
    parent = os.getppid()
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        libc.prctl(1, int(signal.SIGTERM), 0, 0, 0)       # PR_SET_PDEATHSIG = 1
    except (OSError, AttributeError):
        return
    if os.getppid() != parent:                 # the parent died before prctl took effect
        raise SystemExit(143)


def _join_parts(final_path, parts):
    """04_filter from its parts (named by (block, rank) = query order): rename a single part, else
    concatenate into tmp + rename; the parts are removed."""
    import shutil
    parts = sorted(p for p in parts if ".tmp." not in p)
    if len(parts) == 1:
        os.replace(parts[0], final_path)
        return
    tmp = f"{final_path}.tmp.{os.getpid()}"
    with open(tmp, "wb") as out:
        for p in parts:
            with open(p, "rb") as f:
                shutil.copyfileobj(f, out, 16 << 20)
    os.replace(tmp, final_path)
    for p in parts:
        os.unlink(p)


def _wait_for_file(path, timeout_s: float):
    """Contents of `path` once it exists; raises after timeout_s (rank 0 never published it)."""
    import time
    deadline = time.time() + timeout_s
    while not os.path.exists(path):
        if time.time() > deadline:
            raise TimeoutError(f"NCCL id file {path} did not appear within {timeout_s:.0f} s "
                               "(rank 0 failed before publishing it?)")
        time.sleep(0.05)
    with open(path, "rb") as f:
        return f.read()


class _Timing:
    """Wall-clock seconds per phase of a match-db run (main thread), written as JSON on request."""

    def __init__(self):
        import time
        self._now = time.perf_counter
        self.t = {}
        self._t0 = self._now()

    def add(self, key, dt):
        self.t[key] = self.t.get(key, 0.0) + dt

    class _Span:
        def __init__(self, owner, key):
            self.o, self.k = owner, key

        def __enter__(self):
            self.t0 = self.o._now()

        def __exit__(self, *exc):
            self.o.add(self.k, self.o._now() - self.t0)

    def span(self, key):
        return _Timing._Span(self, key)

    def total(self):
        return self._now() - self._t0


def _write_benchmark_log(path, command, wall_s, extra_cols):
    """One log in the format family of scripts/benchmark.py:33-46,64-67: a `# Benchmarking command:`
    line, a tab-separated header and one value line.  The 8 reference columns come first (those
    /usr/bin/time would measure per process are n/a for a batch that shared one resident job);
    the match-stage columns follow."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    head = ["real(s)", "sys(s)", "user(s)", "percent_CPU", "max_RAM(kb)", "FS_inputs", "FS_outputs",
            "elapsed_time_alt(s)"] + [k for k, _ in extra_cols]
    vals = [f"{wall_s:.3f}", "NA", "NA", "NA", "NA", "NA", "NA", f"{wall_s:.3f}"] + [str(v) for _, v in extra_cols]
    with open(path, "w") as f:
        f.write(f"# Benchmarking command: {command}\n" + "\t".join(head) + "\n" + "\t".join(vals) + "\n")


def cmd_match_db(a):
    import json
    import time
    from concurrent.futures import ThreadPoolExecutor as _TPE
    from .matcher import Matcher, nccl_unique_id
    from .match_files import MatchFileSet
    from .cobs_text import _cat
    tm = _Timing()
    rank = int(os.environ.get("PHYLIGN_RANK", -1))
    world = int(os.environ.get("PHYLIGN_WORLD", 1))
    if a.gpus > 1 and rank < 0:
        return _spawn_match_db_workers(a)
    nccl = a.gpus > 1 and rank >= 0          # worker of a multi-GPU job: candidate lists meet over NCCL
    if nccl:                                 # the workers share the host's cores
        _die_with_parent()
        a.load_workers = max(2, a.load_workers // world)
        if not a.write_threads:
            from .match_files import default_threads
            a.write_threads = max(2, default_threads() // world)
    if nccl and a.shard:
        _die("--gpus and --shard are alternatives")
    with open(a.batches) as f:
        batches = sorted(filter(len, map(str.strip, f)))      # Snakefile:32-34
    # the query file is read by a helper thread (native reader, GIL released) while the indexes load
    from concurrent.futures import ThreadPoolExecutor as _QTPE
    _qex = _QTPE(max_workers=1)

    def _read_queries():
        t0 = time.perf_counter()
        q = fasta.QueryFile(a.q)                              # flat arrays: no per-record Python objects
        return q, time.perf_counter() - t0
    qf_future = _qex.submit(_read_queries)
    q_bytes_bound = os.path.getsize(a.q) * (8 if str(a.q).endswith(".gz") else 1)
    qfile = a.qfile or os.path.splitext(os.path.basename(a.q))[0]
    os.makedirs(a.match_dir, exist_ok=True)
    sizes = {}
    if a.index_sizes_table:
        with open(a.index_sizes_table) as f:
            for line in f:
                p = line.split()
                if len(p) >= 2:
                    sizes[os.path.basename(p[0]).replace(".cobs_classic.xz", "")] = int(p[1])
    from . import sharding
    from .cobs_index import IndexStream

    def path_of(b):
        """Where batch b is read from: the decompressed copy of the reference's `decompress_cobs` rule
        (Snakefile:364-387, {decompression_dir}/{batch}.cobs_classic) when there is one, else the .xz."""
        if a.decompression_dir:
            p = os.path.join(a.decompression_dir, f"{b}.cobs_classic")
            if os.path.exists(p):
                return p
        p = os.path.join(a.cobs_dir, f"{b}.cobs_classic.xz")
        return p if os.path.exists(p) else p[:-3]

    def keep_path_of(b):
        """Decompressed copy to leave behind while streaming the .xz (config.yaml keep_cobs_indexes)."""
        if a.keep_cobs_indexes and a.decompression_dir and path_of(b).endswith(".xz"):
            return os.path.join(a.decompression_dir, f"{b}.cobs_classic")
        return None

    merged_inputs, todo = [], []
    for b in batches:
        out = os.path.join(a.match_dir, f"{b}____{qfile}.gz")
        if a.resume and os.path.exists(out):                  # file-granular resume like Snakemake
            merged_inputs.append(out)
        else:
            todo.append(b)
    shapes, headers_of = {}, {}
    with tm.span("plan_s"):
        for b in todo:      # shapes come from the index headers (the first bytes of each stream)
            if not os.path.exists(path_of(b)):
                _die(f"index of batch {b} not found under {a.cobs_dir}")
            st = IndexStream(path_of(b))
            h = st.header
            st.abort()
            headers_of[b] = h
            shapes[b] = sharding.Batch(b, h.n_docs, h.signature_size)
            if b in sizes and sizes[b] != h.header_size + h.body_size:
                _die(f"{b}: decompressed size {h.header_size + h.body_size} != table {sizes[b]}")
    n_shards, shard = 1, 0
    if a.shard:
        shard, n_shards = (int(x) for x in a.shard.split("/"))
    if nccl:
        shard, n_shards = rank, world
    t_ctx = time.perf_counter()
    with Matcher(rank if nccl else a.device, a.hbm_budget) as m:
        tm.add("ctx_create_s", time.perf_counter() - t_ctx)
        # HBM left for indexes = the context's budget minus the working set of one query block
        # (hashes 8 B per base, sequence 1 B, unit tables, result/merge buffers)
        block_bases = min(q_bytes_bound, a.query_block_bases)      # (file size bounds the number of bases)
        working = 10 * block_bases + 24 * (q_bytes_bound // 32 + 1) + (2 << 30)
        free_for_indexes = max(0, m.budget_bytes() - working)
        budget = min(a.round_bytes, free_for_indexes) if a.round_bytes else int(free_for_indexes * 0.95)
        try:
            plan = sharding.assign([shapes[b] for b in todo], n_shards, budget) if todo else sharding.Plan(n_shards)
        except ValueError as e:
            _die(f"{e} (index budget {budget} B = HBM budget {m.budget_bytes()} B minus {working} B for a query "
                 f"block of {block_bases} bases: lower --query-block-bases or raise --hbm-budget)")
        overlap = a.overlap_rounds and len(plan.rounds) > 1
        if overlap:                                           # two rounds resident at once: halve the round size
            try:
                plan = sharding.assign([shapes[b] for b in todo], n_shards, budget // 2)
            except ValueError:                                # a batch needs more than half: one round at a time
                overlap = False
        # 04_filter is merged from the device results of every round (no re-parsing of what was just
        # written); only match files that already existed (--resume) are parsed
        want_filter = bool(a.filter_out) and (n_shards == 1 or nccl)
        if a.filter_out and not want_filter:
            _die("--filter-out needs all batches: use --gpus N, or run `filter` over the match files of all shards")
        brank = sharding.global_batch_ranks(batches)
        rounds = [sorted(x.name for x in rnd[shard]) for rnd in plan.rounds]
        loader = _TPE(max_workers=1)                           # index loads run beside everything else

        def load_round(names):
            t0 = time.perf_counter()
            ids = m.load_indexes([path_of(b) for b in names], names, workers=a.load_workers,
                                 keep_paths=[keep_path_of(b) for b in names], active=False) if names else []
            return ids, time.perf_counter() - t0

        pending_load = loader.submit(load_round, rounds[0]) if rounds else None
        t0 = time.perf_counter()
        qf, read_s = qf_future.result()
        tm.add("read_queries_s", read_s)                      # (thread time; hidden behind the index load)
        tm.add("read_queries_wait_s", time.perf_counter() - t0)
        total_bases = qf.total_bases
        # the merge works on the query dict of filter_queries.py:163-176 (readfq names, duplicates
        # collapsed).  For plain FASTA with unique names that is the record list itself.
        qnames = queries = qid = rec2qid = None
        identity = False
        # multi-GPU: with plain FASTA (identity below) every rank finalises and writes its own slice of the
        # queries ("merge_mode" 1, parts joined by the parent); otherwise rank 0 collects everything
        collect = want_filter
        if collect:
            with tm.span("query_names_s"):
                qnames = qf.names()
                identity = qf.simple and len(set(qnames)) == qf.n
                if not identity:
                    queries, qid = _load_filter_queries(a.q)
                    rec2qid = np.array([qid[nm] for nm in qnames], dtype=np.int64)
                else:
                    qid = {nm: i for i, nm in enumerate(qnames)} if merged_inputs else None
        sharded_merge = bool(nccl and want_filter and identity and not a.bucket_dir)
        if nccl and not sharded_merge:
            collect = want_filter and rank == 0
        n_merge_queries = qf.n if identity else (len(queries) if queries is not None else 0)
        blocks = qf.block_ranges(a.query_block_bases)
        # one resident round + plain FASTA: the device merge of a query block is already the final answer
        # for its queries, so every block's slice of 04_filter is written as a part right away (by the
        # writer thread) and the parts are joined at the end -- no second merge over all blocks
        stream_filter = bool(collect and identity and len(plan.rounds) <= 1 and not merged_inputs and
                             not a.bucket_dir and (not nccl or sharded_merge))
        part_paths = []
        if a.sanitize_queries:      # raw queries: rule fix_query's base transform on the device, once; the
            with tm.span("sanitize_queries_s"):   # 04_filter FASTA then carries the sanitised sequences too
                m.fix_bases(qf.seqs[:qf.total_bases])

        def own_range(q0, q1):
            """Queries of block [q0, q1) whose merged lists this rank holds (phy_merged_range)."""
            if not sharded_merge:
                return q0, q1
            n = q1 - q0
            return q0 + n * rank // world, q0 + n * (rank + 1) // world
        pieces, refs_by_rank = [], {}
        if collect and merged_inputs:
            with tm.span("parse_existing_s"):
                pieces, refs_by_rank = _parsed_pieces(merged_inputs, qid, brank, open(os.devnull, "w"))
            if sharded_merge:                                 # keep the candidates of this rank's queries only
                own = np.zeros(qf.n, dtype=bool)
                for q0, q1 in blocks:
                    lo, hi = own_range(q0, q1)
                    own[lo:hi] = True
                pieces = [(qs[own[qs]], cs[own[qs]]) for qs, cs in pieces]
        if nccl:                                              # rank 0 publishes the NCCL id through a file
            id_file = os.environ["PHYLIGN_NCCL_ID_FILE"]
            if rank == 0:
                with open(id_file + ".tmp", "wb") as f:
                    f.write(nccl_unique_id())
                os.replace(id_file + ".tmp", id_file)
            m.nccl_init(_wait_for_file(id_file, float(os.environ.get("PHYLIGN_NCCL_ID_TIMEOUT", 300))), rank, world)
            m.set_option("shard_query_upload", 1)             # every worker passes the same queries
            m.set_option("merge_mode", int(sharded_merge))
        if stream_filter:                                     # accessions of every batch, from the index headers
            for b in todo:
                refs_by_rank[brank[b]] = [_ref_of(n) for n in headers_of[b].doc_names]
            os.makedirs(os.path.dirname(os.path.abspath(a.filter_out)), exist_ok=True)
        # page-locked result buffers pay off when they are reused block after block; a single
        # (round, block) run fetches once, so plain host memory is cheaper than pinning it
        m.set_option("pinned_results", int(len(blocks) * max(1, len(plan.rounds)) > 8))   # (2 pool generations to pin)
        wstats, gpu_phase_ms, gathered_total, n_writer_blocks = [], np.zeros(3), 0, 0
        direct_merged = direct_arrays = None                   # set when one device merge is already the final answer
        bg = _TPE(max_workers=1)                               # the writer thread (format + gzip + append)
        try:
            for ri, mine in enumerate(rounds):                 # resident round: load, match, write, evict
                if pending_load is None:
                    pending_load = loader.submit(load_round, mine)
                t0 = time.perf_counter()
                loaded, load_s = pending_load.result()
                tm.add("index_load_s", load_s)
                tm.add("index_load_wait_s", time.perf_counter() - t0)
                pending_load = None
                if overlap and ri + 1 < len(rounds):           # decode + push round r+1 while r is matched
                    pending_load = loader.submit(load_round, rounds[ri + 1])
                if not mine and not (nccl and want_filter):
                    continue                                   # (under NCCL every rank joins every merge)
                with tm.span("set_ranks_s"):
                    m.set_ranks(batches)
                    m.set_active_only(loaded)
                fs = MatchFileSet({idx: os.path.join(a.match_dir, f"{m.indexes[idx].batch}____{qfile}.gz")
                                   for idx in loaded}, m.indexes, gzip_level=1, threads=a.write_threads)
                n_hit_queries = {idx: 0 for idx in loaded}
                round_t0 = time.perf_counter()
                round_gpu_ms, round_bytes_by_idx = 0.0, {idx: 0 for idx in loaded}
                fut = None
                try:
                    for bi, (q0, q1) in enumerate(blocks):
                        if len(blocks) > 1 or ri == 0:          # one block: queries stay resident across rounds
                            with tm.span("set_queries_s"):
                                m.set_queries_raw(qf.seqs, qf.soffs[q0:q1 + 1])
                        with tm.span("gpu_match_s"):
                            m.match_run(a.t, top_n=a.n, floor_mode=a.floor, merge_top_n=a.n if want_filter else 0)
                        ph = m.phase_ms()
                        gpu_phase_ms += ph[:3]
                        round_gpu_ms += float(sum(ph[:3]))
                        gathered_total += m.gathered_bytes()
                        for idx in loaded:
                            round_bytes_by_idx[idx] += m.gathered_bytes_of(idx)
                        with tm.span("fetch_results_s"):
                            res = m.fetch()
                        for idx in loaded:
                            n_hit_queries[idx] += len(res.units_of(idx))
                        if fut is not None:                     # at most one block in flight behind the GPU
                            with tm.span("writer_wait_s"):
                                fut.result()
                        fut = bg.submit(lambda r=res, ho=qf.hoffs[q0:q1 + 1]: fs.write_block(qf.headers, ho, r._owner.ptr))
                        n_writer_blocks += 1
                        if collect:                             # this block's top-N + ties per query and round
                            with tm.span("fetch_merged_s"):
                                moffs, mc = m.merged()
                            if stream_filter:
                                lo, hi = own_range(q0, q1)
                                # one GPU: the blocks are appended in order to one temporary file; several GPUs:
                                # one part per (block, rank), joined by the parent
                                part = (f"{a.filter_out}.tmp.{os.getpid()}" if not nccl else
                                        f"{a.filter_out}.part.{bi:06d}.{rank:04d}")
                                if part not in part_paths:
                                    part_paths.append(part)
                                    if not nccl and os.path.exists(part):
                                        os.unlink(part)
                                from .cobs_text import write_filter_fasta_native
                                prev = fut
                                fut = bg.submit(lambda pv=prev, pa=part, ow=m._merged_owner, l=lo - q0, h=hi - q0, qb=q0:
                                                (pv.result() if pv is not None else None,
                                                 write_filter_fasta_native(pa, ow.ptr, qf, refs_by_rank, l, h, qb,
                                                                           append=not nccl)))
                            elif identity and len(rounds) == 1 and len(blocks) == 1 and not pieces:
                                direct_merged = m._merged_owner   # already the global answer: no host re-merge
                                direct_arrays = (moffs, mc)
                            else:
                                q_of = q0 + np.repeat(np.arange(q1 - q0, dtype=np.int64),
                                                      np.diff(moffs.astype(np.int64)))
                                pieces.append((q_of if identity else rec2qid[q_of], np.array(mc)))
                            if sharded_merge:
                                lo, hi = m.merged_range()
                                assert (q0 + lo, q0 + hi) == own_range(q0, q1)
                        elif nccl and want_filter:
                            m.merged()                          # non-holders still take part in the fetch
                    if fut is not None:
                        with tm.span("writer_wait_s"):
                            fut.result()
                    with tm.span("commit_files_s"):
                        fs.commit()
                except BaseException:
                    if fut is not None:
                        try:
                            fut.result()
                        except Exception:
                            pass
                    fs.abort()
                    for pp in part_paths:                       # nothing partial may stay behind
                        for cand in (pp, f"{pp}.tmp.{os.getpid()}"):
                            if os.path.exists(cand):
                                os.unlink(cand)
                    raise
                wstats.append(fs.stats_dict())
                round_wall = time.perf_counter() - round_t0
                kmers = int(np.maximum(np.diff(qf.soffs.astype(np.int64)) - 30, 0).sum())
                round_alg = sum((m.indexes[i].header.n_docs + 7) // 8 for i in loaded) or 1
                for idx in loaded:
                    ix = m.indexes[idx]
                    print(f"[match-db] {ix.batch}: {n_hit_queries[idx]} queries with hits", file=sys.stderr)
                    refs_by_rank[ix.batch_rank] = [_ref_of(n) for n in ix.doc_names]
                    if a.benchmark_dir:                         # logs/benchmarks/run_cobs/{batch}____{qfile}.txt
                        rb = (ix.header.n_docs + 7) // 8
                        share = rb / round_alg                  # of the round (one fused launch per row class)
                        wall = round_wall * share
                        _write_benchmark_log(
                            os.path.join(a.benchmark_dir, f"{ix.batch}____{qfile}.txt"),
                            f"phylign_b200 match-db (resident round {ri}, {len(loaded)} batches on GPU "
                            f"{m.device}; per-batch share by row bytes)", wall,
                            [("batch", ix.batch), ("qfile", qfile), ("gpu_ms", f"{round_gpu_ms * share:.3f}"),
                             ("bases/s", f"{total_bases / max(wall, 1e-9):.4g}"),
                             ("kmer_docs/s", f"{kmers * ix.header.n_docs / max(wall, 1e-9):.4g}"),
                             ("GB/s_gathered", f"{round_bytes_by_idx[idx] / max(round_gpu_ms * share, 1e-9) / 1e6:.1f}"),
                             ("rows_read/all", f"{round_bytes_by_idx[idx] / max(1, kmers * rb):.4f}"),
                             ("match_file_bytes", fs.file_bytes.get(idx, 0))])
                if ri + 1 < len(rounds):                        # (after the last round the process ends anyway)
                    with tm.span("evict_s"):
                        for idx in loaded:
                            m.evict(idx)
        finally:
            if pending_load is not None:
                try:
                    pending_load.result()
                except Exception:
                    pass
            bg.shutdown(wait=True)
            loader.shutdown(wait=True)
        if collect and nccl:                                  # accessions of the batches other ranks hold
            for b in todo:
                if brank[b] not in refs_by_rank:
                    refs_by_rank[brank[b]] = [_ref_of(n) for n in headers_of[b].doc_names]
        if stream_filter:
            if not nccl:                                      # (multi-GPU: the parent joins the workers' parts)
                with tm.span("write_filter_s"):
                    if part_paths:
                        os.replace(part_paths[0], a.filter_out)
                    else:
                        _atomic_write(a.filter_out, b"", gz=False)
        elif collect:
            os.makedirs(os.path.dirname(os.path.abspath(a.filter_out)), exist_ok=True)
            with tm.span("final_merge_s"):
                if direct_merged is None:
                    direct_arrays = _merge_pieces(m, n_merge_queries, pieces, a.n)
                    direct_merged = m._merged_owner
            with tm.span("write_filter_s"):
                if sharded_merge:                             # this rank's slice of every block, one part each
                    from .cobs_text import write_filter_fasta_native
                    for bi, (q0, q1) in enumerate(blocks):
                        lo, hi = own_range(q0, q1)
                        write_filter_fasta_native(f"{a.filter_out}.part.{bi:06d}.{rank:04d}", direct_merged.ptr, qf,
                                                  refs_by_rank, lo, hi)
                elif identity:                                # flat arrays straight into the file (tmp + rename)
                    from .cobs_text import write_filter_fasta_native
                    write_filter_fasta_native(a.filter_out, direct_merged.ptr, qf, refs_by_rank)
                else:
                    _atomic_write(a.filter_out, format_filter_fasta_fast(list(queries.items()), direct_merged.ptr,
                                                                         refs_by_rank), gz=False)
            if a.bucket_dir:      # per-batch "reference -> queries to align" tables for stage 05
                from .cobs_text import candidate_buckets, format_bucket_tsv
                with tm.span("buckets_s"):
                    moffs, mc = direct_arrays
                    buckets = candidate_buckets(qnames if identity else list(queries), moffs, mc, refs_by_rank)
                    os.makedirs(a.bucket_dir, exist_ok=True)
                    for b in batches:
                        _atomic_write(os.path.join(a.bucket_dir, f"{b}____{qfile}.candidates.tsv"),
                                      format_bucket_tsv(buckets.get(brank[b], [])).encode(), gz=False)
        if nccl:
            m.nccl_finalize()    # every worker gets here after the same number of collectives
        m.release_at_exit()      # HBM goes back with the process: no cudaFree per index on the way out
        if a.timing_json and (not nccl or rank == 0):
            w = {k: sum(d[k] for d in wstats) for k in (wstats[0] if wstats else {}) if k != "threads"}
            out = {"total_s": tm.total(), "phases_s": {k: round(v, 4) for k, v in tm.t.items()},
                   "writer": dict(w, threads=wstats[0]["threads"] if wstats else 0, blocks=n_writer_blocks),
                   "gpu_phase_ms_hash_gather_merge": [round(float(x), 3) for x in gpu_phase_ms],
                   "gathered_bytes": int(gathered_total), "rounds": len(rounds), "query_blocks": len(blocks),
                   "overlap_rounds": bool(overlap), "n_queries": qf.n, "bases": int(total_bases),
                   "n_batches": len(todo), "rank": max(rank, 0), "world": world,
                   "direct_device_merge": bool(stream_filter or (identity and len(rounds) == 1 and len(blocks) == 1)),
                   "filter_written_per_block": bool(stream_filter),
                   "native_filter_writer": bool(identity)}
            tmp = a.timing_json + f".tmp.{os.getpid()}"
            with open(tmp, "w") as f:
                json.dump(out, f)
            os.replace(tmp, a.timing_json)


# ------------------------------------------------------------------------------------ argparse
def build_parser():
    ap = argparse.ArgumentParser(prog="phylign_b200", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)

    cobs = sub.add_parser("cobs", help="cobs-compatible front end").add_subparsers(dest="cobs_cmd", required=True)
    q = cobs.add_parser("query")
    q.add_argument("--load-complete", action="store_true", help="accepted for compatibility (always resident in HBM)")
    q.add_argument("-t", type=float, default=0.8, help="k-mer threshold (cobs default 0.8)")
    q.add_argument("-T", type=int, default=0, help="threads: accepted and ignored")
    q.add_argument("-i", required=True, help="index file (.cobs_classic, .cobs_classic.xz or a pipe)")
    q.add_argument("--index-sizes", type=int, default=None, help="decompressed index size in bytes (verified)")
    q.add_argument("-f", required=True, help="query FASTA")
    q.add_argument("--top-n", type=int, default=0, help="fuse postprocess_cobs.py -n N (names get the '_acc' form)")
    q.add_argument("--floor", action="store_true", help="threshold = floor(t*K) instead of ceil (SURVEY A.6)")
    q.add_argument("--device", type=int, default=0)
    q.add_argument("--query-block-bases", type=int, default=2 * 10 ** 9,
                   help="process the queries in blocks of at most this many bases (HBM for the hashes)")
    q.add_argument("--server", default=None, help="socket of a resident `serve` process (or $PHYLIGN_SERVER)")
    q.add_argument("--sanitize-queries", action="store_true",
                   help="queries must be upper-case ACGT (the contract of intermediate/01_queries_merged, "
                        "Snakefile:326-332); any other letter ends the run with an error, like a cobs abort.  With this "
                        "flag the bases are upper-cased and every other letter becomes A on the GPU first")
    q.set_defaults(fn=cmd_cobs_query)

    r = sub.add_parser("run-cobs-streaming")
    r.add_argument("kmer_thres", type=float)
    r.add_argument("threads")
    r.add_argument("cobs_index_xz")
    r.add_argument("uncompressed_size", type=int)
    r.add_argument("query")
    r.add_argument("--device", type=int, default=0)
    r.set_defaults(fn=lambda a: cmd_cobs_query(argparse.Namespace(
        t=a.kmer_thres, T=0, i=a.cobs_index_xz, index_sizes=a.uncompressed_size, f=a.query, top_n=0,
        floor=False, device=a.device, query_block_bases=2 * 10 ** 9, sanitize_queries=False)))

    sv = sub.add_parser("serve", help="keep indexes resident in HBM and answer `cobs query --server` requests")
    sv.add_argument("--socket", required=True)
    sv.add_argument("--device", type=int, default=0)
    sv.add_argument("--hbm-budget", type=int, default=0)
    sv.add_argument("--preload", nargs="*", default=[], help="index files to load before serving")
    sv.set_defaults(fn=lambda a: __import__("phylign_b200.server", fromlist=["serve"]).serve(
        a.socket, a.device, a.hbm_budget, a.preload))

    p = sub.add_parser("postprocess")
    p.add_argument("-n", dest="keep", required=True, type=int, metavar="int", help="no. of best hits to keep")
    p.set_defaults(fn=lambda a: _postprocess_stream(sys.stdin, sys.stdout, a.keep))

    f = sub.add_parser("filter")
    f.add_argument("match_fn", nargs="+")
    f.add_argument("-q", dest="query_fn", required=True, metavar="str", help="query file")
    f.add_argument("-n", dest="keep", type=int, default=100, metavar="int", help="no. of best hits to keep [100]")
    f.add_argument("--device", type=int, default=0)
    f.set_defaults(fn=cmd_filter)

    fq = sub.add_parser("fix-query", help="seqtk seq -A -U -C | awk non-ACGT->A (Snakefile:326-332), all inputs concatenated")
    fq.add_argument("inputs", nargs="+")
    fq.add_argument("--device", type=int, default=None,
                    help="run the base transform on this GPU (phy_fix_bases) instead of the host table")
    fq.set_defaults(fn=cmd_fix_query)

    d = sub.add_parser("match-db")
    d.add_argument("--cobs-dir", required=True)
    d.add_argument("--batches", required=True, help="file with one batch name per line (config.yaml: batches)")
    d.add_argument("-q", required=True, help="merged query FASTA (intermediate/01_queries_merged/{qfile}.fa)")
    d.add_argument("--qfile", default=None, help="qfile wildcard (default: stem of -q)")
    d.add_argument("--match-dir", default="intermediate/03_match")
    d.add_argument("--filter-out", default=None, help="intermediate/04_filter/{qfile}.fa")
    d.add_argument("-t", type=float, default=0.7, help="cobs_kmer_thres")
    d.add_argument("-n", type=int, default=100, help="nb_best_hits")
    d.add_argument("--floor", action="store_true")
    d.add_argument("--index-sizes-table", default=None, help="data/decompressed_indexes_sizes.txt (verified)")
    d.add_argument("--resume", action="store_true", help="skip batches whose match file exists")
    d.add_argument("--hbm-budget", type=int, default=0)
    d.add_argument("--shard", default=None, metavar="I/N",
                   help="process only the batches the LPT plan gives to GPU I of N (one process per GPU); "
                        "run `filter` over all match files afterwards")
    d.add_argument("--load-workers", type=int, default=8, help="concurrent xz decoders while loading")
    d.add_argument("--query-block-bases", type=int, default=2 * 10 ** 9,
                   help="process the queries in blocks of at most this many bases (HBM for the hashes)")
    d.add_argument("--gpus", type=int, default=1,
                   help="N > 1: one worker process per GPU (devices 0..N-1), batches placed by the LPT plan, "
                        "per-GPU candidate lists merged over NCCL on rank 0 (writes --filter-out directly)")
    d.add_argument("--bucket-dir", default=None,
                   help="also write {batch}____{qfile}.candidates.tsv (reference -> queries), the mapping "
                        "batch_align.py:126-171 derives per batch from the 04_filter FASTA")
    d.add_argument("--round-bytes", type=int, default=0,
                   help="HBM bytes of indexes resident at once (default 90%% of --hbm-budget, else 160e9); "
                        "batches beyond it are streamed through in further rounds")
    d.add_argument("--device", type=int, default=0)
    d.add_argument("--write-threads", type=int, default=0,
                   help="host threads that format + gzip the match files (default: all cores)")
    d.add_argument("--decompression-dir", default=None,
                   help="{decompression_dir} of the reference (config.yaml): a {batch}.cobs_classic found there is "
                        "loaded instead of the .xz (no LZMA decode)")
    d.add_argument("--keep-cobs-indexes", action="store_true",
                   help="leave the decompressed {batch}.cobs_classic in --decompression-dir while streaming the .xz "
                        "(config.yaml keep_cobs_indexes; same bytes as rule decompress_cobs writes)")
    d.add_argument("--no-overlap-rounds", dest="overlap_rounds", action="store_false",
                   help="when the batches need several resident rounds: do not load round r+1 while round r is "
                        "matched (default: overlap, each round then uses half of the index budget)")
    d.add_argument("--benchmark-dir", default=None,
                   help="write logs/benchmarks/run_cobs-style {batch}____{qfile}.txt files (scripts/benchmark.py format)")
    d.add_argument("--timing-json", default=None, help="write the wall-clock breakdown of the run here")
    d.add_argument("--sanitize-queries", action="store_true",
                   help="the query FASTA has not gone through rule fix_query: upper-case its bases and turn every "
                        "letter outside ACGT into A on the GPU before matching (Snakefile:326-332)")
    d.set_defaults(fn=cmd_match_db)
    return ap


def main(argv=None):
    a = build_parser().parse_args(argv)
    try:
        a.fn(a)
    except SystemExit:
        raise
    except BrokenPipeError:
        sys.exit(1)
    except Exception as e:           # one line on stderr, non-zero exit (set -euo pipefail callers)
        _die(f"{type(e).__name__}: {e}")


if __name__ == "__main__":
    main()
