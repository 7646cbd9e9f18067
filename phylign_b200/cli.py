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
from .cobs_text import format_cobs_text_fast, format_filter_fasta_fast


def _die(msg, code=1):
    print(f"phylign_b200: error: {msg}", file=sys.stderr)
    sys.exit(code)


def _postprocess_stream(fin, fout, keep: int):
    """Text filter with the exact semantics of postprocess_cobs.py:21-38."""
    i, min_kmers = 0, 0
    for x in fin:
        if x[0] == "*":
            i, min_kmers = 0, 0
            fout.write(x)
            continue
        y = "_" + x.partition("_")[2]
        i += 1
        if i < keep:
            fout.write(y)
        elif i == keep:
            fout.write(y)
            min_kmers = int(y.split("\t")[-1])
        elif int(y.split("\t")[-1]) == min_kmers:
            fout.write(y)


def query_blocks(records, max_bases: int):
    """Split the query list into consecutive blocks of at most max_bases bases (at least one
    record each): the k-mer hashes of a block (8 B per base) must fit in HBM beside the indexes."""
    start, acc = 0, 0
    for i, (_, seq) in enumerate(records):
        if i > start and acc + len(seq) > max_bases:
            yield start, records[start:i]
            start, acc = i, 0
        acc += len(seq)
    if start < len(records) or not records:
        yield start, records[start:]


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
    records = fasta.read_cobs_records(a.f)
    with Matcher(a.device) as m:
        idx = m.load_index(a.i, batch="index")
        hdr = m.indexes[idx].header
        if a.index_sizes is not None and a.index_sizes != hdr.header_size + hdr.body_size:
            _die(f"--index-sizes {a.index_sizes} != header {hdr.header_size} + body {hdr.body_size}")
        for _, block in query_blocks(records, a.query_block_bases):
            m.set_queries(block)
            res = m.match(a.t, top_n=a.top_n, floor_mode=a.floor)
            sys.stdout.buffer.write(format_cobs_text_fast(block, res, m.indexes[idx], strip_prefix=a.top_n > 0))
    sys.stdout.buffer.flush()


# ------------------------------------------------------------------------------------ filter
def parse_match_file(path):
    """[(qname, [(ref, kmers)])] with the parsing rules of filter_queries.py:27-66."""
    blocks = []
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as f:
        for x in f:
            x = x.strip()
            if not x:
                continue
            if x[0] == "*":
                parts = x[1:].split("\t")
                int(parts[1])
                blocks.append((parts[0].split(" ")[0], []))
            else:
                if not blocks:
                    raise ValueError(f"{path}: hit line before any query header")
                tmp_name, kmers = x.split()
                _rid, ref = tmp_name.split("_")       # exactly one underscore (filter_queries.py:64)
                blocks[-1][1].append((ref, int(kmers)))
    if not blocks:
        raise ValueError(f"{path}: empty match file")
    return blocks


def parse_match_file_native(path):
    """Same content as parse_match_file, produced by the library's C++ parser, as arrays:
    (qnames [str per block], first_hit uint64[n_blocks+1], ref_ids int64[n_hits], refs_sorted [str],
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
    qnames = [text[o:o + n].decode() for o, n in zip(q_off.tolist(), q_len.tolist())]
    if nh:
        width = int(ref_len.max())
        buf = np.frombuffer(text, dtype=np.uint8)
        cols = np.arange(width, dtype=np.int64)
        idx = np.minimum(ref_off[:, None] + cols[None, :], len(buf) - 1)
        mat = np.where(cols[None, :] < ref_len[:, None], buf[idx], 0).astype(np.uint8)
        names = np.ascontiguousarray(mat).view(f"S{width}").ravel()     # NUL padded fixed-width strings
        uniq, inv = np.unique(names, return_inverse=True)
        refs_sorted = [u.decode() for u in uniq.tolist()]
        ref_ids = inv.astype(np.int64)
    else:
        refs_sorted, ref_ids = [], np.zeros(0, np.int64)
    return qnames, first_hit, ref_ids, refs_sorted, kmers


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
    for fn in match_fns:
        batch = os.path.basename(fn).split("____")[0]
        print(f"Translating matches {fn}", file=log)
        by_batch.setdefault(batch, []).append(parse_match_file_native(fn))
    pieces, refs_by_rank = [], {}
    for batch, parsed in by_batch.items():
        br = brank[batch]
        refs = sorted({r for p in parsed for r in p[3]})          # accessions of the batch, str order
        rr = {r: i for i, r in enumerate(refs)}
        refs_by_rank[br] = refs
        for qnames, first_hit, ref_ids, refs_sorted, kmers in parsed:
            try:
                q_of_block = np.array([qid[q] for q in qnames], dtype=np.int64)
            except KeyError as e:
                raise KeyError(f"query {e.args[0]!r} of batch {batch} is not in the query file") from None
            remap = np.array([rr[r] for r in refs_sorted], dtype=np.int64)   # file-local id -> batch rank
            rank = remap[ref_ids] if len(ref_ids) else np.zeros(0, np.int64)
            c = np.zeros(len(kmers), dtype=CAND_DT)
            c["score"], c["batch_rank"], c["doc"], c["ref_rank"] = kmers, br, rank, rank
            counts = np.diff(first_hit.astype(np.int64))
            pieces.append((np.repeat(q_of_block, counts), c))
    return pieces, refs_by_rank


def _final_merge(m, queries, pieces, refs_by_rank, keep: int) -> str:
    """Global top-N + ties over candidate pieces (filter_queries.py:123-150) on the GPU."""
    from .matcher import CAND_DT
    nq = len(queries)
    qs = np.concatenate([p[0] for p in pieces]) if pieces else np.zeros(0, np.int64)
    cs = np.concatenate([p[1] for p in pieces]) if pieces else np.zeros(0, CAND_DT)
    order = np.argsort(qs, kind="stable")
    offs = np.zeros(nq + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(np.bincount(qs, minlength=nq), dtype=np.uint64)
    m.merge_host(offs, cs[order], keep)
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
    """`match-db --gpus N`: one worker process per GPU (device = rank), NCCL merge on rank 0."""
    import subprocess
    import tempfile
    id_file = os.path.join(tempfile.mkdtemp(prefix="phylign_nccl_"), "id")
    argv = [x for x in sys.argv[1:]]
    procs = []
    for r in range(a.gpus):
        env = dict(os.environ, PHYLIGN_RANK=str(r), PHYLIGN_WORLD=str(a.gpus), PHYLIGN_NCCL_ID_FILE=id_file)
        procs.append(subprocess.Popen([sys.executable, "-m", "phylign_b200.cli"] + argv, env=env))
    rcs = [p.wait() for p in procs]
    if any(rcs):
        _die(f"match-db workers failed (exit codes {rcs})")


def cmd_match_db(a):
    from .matcher import Matcher, nccl_unique_id
    rank = int(os.environ.get("PHYLIGN_RANK", -1))
    world = int(os.environ.get("PHYLIGN_WORLD", 1))
    if a.gpus > 1 and rank < 0:
        return _spawn_match_db_workers(a)
    nccl = a.gpus > 1 and rank >= 0          # worker of a multi-GPU job: candidate lists meet over NCCL
    if nccl and a.shard:
        _die("--gpus and --shard are alternatives")
    with open(a.batches) as f:
        batches = sorted(filter(len, map(str.strip, f)))      # Snakefile:32-34
    records = fasta.read_cobs_records(a.q)
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
        p = os.path.join(a.cobs_dir, f"{b}.cobs_classic.xz")
        return p if os.path.exists(p) else p[:-3]

    merged_inputs, todo = [], []
    for b in batches:
        out = os.path.join(a.match_dir, f"{b}____{qfile}.gz")
        if a.resume and os.path.exists(out):                  # file-granular resume like Snakemake
            merged_inputs.append(out)
        else:
            todo.append(b)
    shapes = {}
    for b in todo:      # shapes come from the index headers (the first bytes of each xz stream)
        if not os.path.exists(path_of(b)):
            _die(f"index of batch {b} not found under {a.cobs_dir}")
        st = IndexStream(path_of(b))
        h = st.header
        st.abort()
        shapes[b] = sharding.Batch(b, h.n_docs, h.signature_size)
        if b in sizes and sizes[b] != h.header_size + h.body_size:
            _die(f"{b}: decompressed size {h.header_size + h.body_size} != table {sizes[b]}")
    n_shards, shard = 1, 0
    if a.shard:
        shard, n_shards = (int(x) for x in a.shard.split("/"))
    if nccl:
        shard, n_shards = rank, world
    budget = a.round_bytes or (int(a.hbm_budget * 0.9) if a.hbm_budget else 160 * 10 ** 9)
    plan = sharding.assign([shapes[b] for b in todo], n_shards, budget) if todo else sharding.Plan(n_shards)
    # 04_filter is merged from the device results of every round (no re-parsing of what was just
    # written); only match files that already existed (--resume) are parsed
    want_filter = bool(a.filter_out) and (n_shards == 1 or nccl)
    if a.filter_out and not want_filter:
        _die("--filter-out needs all batches: use --gpus N, or run `filter` over the match files of all shards")
    collect = want_filter and (not nccl or rank == 0)     # who assembles 04_filter
    queries, qid = _load_filter_queries(a.q) if collect else ({}, {})
    brank = sharding.global_batch_ranks(batches)
    rec2qid = np.array([qid[h.split(" ")[0]] for h, _ in records], dtype=np.int64) if collect else None
    pieces, refs_by_rank = [], {}
    if collect and merged_inputs:
        pieces, refs_by_rank = _parsed_pieces(merged_inputs, qid, brank, open(os.devnull, "w"))
    with Matcher(rank if nccl else a.device, a.hbm_budget) as m:
        if nccl:                                              # rank 0 publishes the NCCL id through a file
            id_file = os.environ["PHYLIGN_NCCL_ID_FILE"]
            if rank == 0:
                with open(id_file + ".tmp", "wb") as f:
                    f.write(nccl_unique_id())
                os.replace(id_file + ".tmp", id_file)
            import time
            for _ in range(6000):
                if os.path.exists(id_file):
                    break
                time.sleep(0.05)
            m.nccl_init(open(id_file, "rb").read(), rank, world)
        for rnd in plan.rounds:                               # resident round: load, match, write, evict
            mine = sorted(x.name for x in rnd[shard])
            if not mine and not (nccl and want_filter):
                continue                                      # (under NCCL every rank joins every merge)
            loaded = m.load_indexes([path_of(b) for b in mine], mine, workers=a.load_workers) if mine else []
            m.set_ranks(batches)
            texts = {idx: [] for idx in loaded}               # per index: one text piece per query block
            n_hit_queries = {idx: 0 for idx in loaded}
            for q0, block in query_blocks(records, a.query_block_bases):
                m.set_queries(block)
                m.match_run(a.t, top_n=a.n, floor_mode=a.floor, merge_top_n=a.n if want_filter else 0)
                res = m.fetch()
                if loaded:
                    format_cobs_text_fast(block, res, m.indexes[loaded[0]], strip_prefix=True)   # warm the header cache
                with ThreadPoolExecutor(max_workers=max(1, a.load_workers)) as ex:   # C++ formatter releases the GIL
                    for idx, text in zip(loaded, ex.map(
                            lambda i: format_cobs_text_fast(block, res, m.indexes[i], strip_prefix=True), loaded)):
                        texts[idx].append(text)
                        n_hit_queries[idx] += len(res.units_of(idx))
                if collect:                                   # this block's top-N + ties per query and round
                    moffs, mc = m.merged()
                    q_of = q0 + np.repeat(np.arange(len(block), dtype=np.int64), np.diff(moffs.astype(np.int64)))
                    pieces.append((rec2qid[q_of], np.array(mc)))

            def write_one(idx):       # gzip (zlib) releases the GIL: one thread per file
                ix = m.indexes[idx]
                _atomic_write(os.path.join(a.match_dir, f"{ix.batch}____{qfile}.gz"), b"".join(texts[idx]), gz=True)
                return idx

            with ThreadPoolExecutor(max_workers=max(1, a.load_workers)) as ex:
                for idx in ex.map(write_one, loaded):
                    ix = m.indexes[idx]
                    print(f"[match-db] {ix.batch}: {n_hit_queries[idx]} queries with hits", file=sys.stderr)
                    refs_by_rank[ix.batch_rank] = [_ref_of(n) for n in ix.doc_names]
            for idx in loaded:
                m.evict(idx)
        if collect and nccl:                                  # accessions of the batches other ranks hold
            for b in todo:
                if brank[b] not in refs_by_rank:
                    st = IndexStream(path_of(b))
                    refs_by_rank[brank[b]] = [_ref_of(n) for n in st.header.doc_names]
                    st.abort()
        if collect:
            fa = _final_merge(m, queries, pieces, refs_by_rank, a.n)
            os.makedirs(os.path.dirname(os.path.abspath(a.filter_out)), exist_ok=True)
            _atomic_write(a.filter_out, fa.encode(), gz=False)
            if a.bucket_dir:      # per-batch "reference -> queries to align" tables for stage 05
                from .cobs_text import candidate_buckets, format_bucket_tsv
                moffs, mc = m.merged()
                buckets = candidate_buckets(list(queries), moffs, mc, refs_by_rank)
                os.makedirs(a.bucket_dir, exist_ok=True)
                for b in batches:
                    _atomic_write(os.path.join(a.bucket_dir, f"{b}____{qfile}.candidates.tsv"),
                                  format_bucket_tsv(buckets.get(brank[b], [])).encode(), gz=False)


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
        floor=False, device=a.device, query_block_bases=2 * 10 ** 9)))

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
    fq.set_defaults(fn=lambda a: sys.stdout.write("".join(fasta.fix_query_file(p) for p in a.inputs)))

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
