"""Host-side driver of the match stage: the Python mirror of what the reference does with
`cobs query | postprocess_cobs.py` per batch and `filter_queries.py` over all batches
(/root/reference/Snakefile:390-520), on top of the C ABI of libphylign_cuda.so.

Python keeps what SURVEY.md 8(b) leaves to the host: header parsing, document names,
rank tables for the merge key, text formatting.  All arithmetic runs in the CUDA library;
nothing here falls back to a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .cobs_index import IndexStream, parse_bytes, ref_of

UNIT_DT = np.dtype([("query", "<u4"), ("index", "<u4"), ("n_pass", "<u4"), ("n_kept", "<u4"),
                    ("offset", "<u8")])
HIT_DT = np.dtype([("doc", "<u4"), ("score", "<u4")])
CAND_DT = np.dtype([("score", "<u4"), ("batch_rank", "<u4"), ("doc", "<u4"), ("ref_rank", "<u4")])


class ResidentIndex:
    def __init__(self, idx_id, batch, header):
        self.idx_id = idx_id
        self.batch = batch
        self.header = header
        self.doc_names = header.doc_names
        self.batch_rank = idx_id


class _Owned:
    """Keeps a library-owned result block alive; frees it when the last view holder dies."""

    def __init__(self, ptr, free_fn):
        self.ptr, self._free = ptr, free_fn

    def __del__(self):
        try:
            if self.ptr:
                self._free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _view(ptr, n, dtype, owner):
    """Zero-copy numpy view of n records at a ctypes pointer (pinned, owned by `owner`).
    The array (and every slice of it) keeps `owner` alive: numpy's base chain ends at the ctypes
    buffer object, which carries a reference to the owner, so the library block is only returned
    to the pinned pool when the last view is gone."""
    if n == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(C.addressof(ptr.contents))
    buf._phy_owner = owner
    a = np.frombuffer(buf, dtype=dtype)
    a.flags.writeable = False
    return a


class MatchResult:
    """Non-empty (query, index) blocks of one match call.

    `units` / `hits` are zero-copy views of the library's pinned result buffers; they stay
    valid as long as this object lives (copy them to keep them longer)."""

    def __init__(self, units, hits, n_kmers, n_queries, h2d_bytes, d2h_bytes, owner=None):
        self.units, self.hits, self.n_kmers = units, hits, n_kmers
        self.n_queries = n_queries
        self.h2d_bytes, self.d2h_bytes = h2d_bytes, d2h_bytes
        self._owner = owner

    def units_of(self, idx_id):
        lo = np.searchsorted(self.units["index"], idx_id, "left")
        hi = np.searchsorted(self.units["index"], idx_id, "right")
        return self.units[lo:hi]

    def hits_of(self, unit):
        o = int(unit["offset"])
        return self.hits[o:o + int(unit["n_kept"])]


class PinnedBuffer:
    """Page-locked host buffer from phy_host_alloc, exposed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        self._L = _lib.load()
        self._p = C.c_void_p()
        _lib.check(self._L.phy_host_alloc(max(1, nbytes), C.byref(self._p)))
        self.nbytes = nbytes
        self.array = np.frombuffer((C.c_char * max(1, nbytes)).from_address(self._p.value), dtype=np.uint8)[:nbytes]

    @property
    def ptr(self):
        return self._p.value

    def __del__(self):
        try:
            if self._p:
                self.array = None
                self._L.phy_host_free(self._p)
                self._p = C.c_void_p()
        except Exception:
            pass


class _SerializedLib:
    """The library as one Matcher sees it: every call holds the Matcher's lock, because a phy_ctx is
    driven by one host thread at a time (include/phylign_cuda.h).  Lets a loader thread push the next
    round's indexes between the match calls of the main thread."""

    def __init__(self, L, lock):
        self._L, self._lock, self._errors = L, lock, {}

    def __getattr__(self, name):
        import threading
        fn = getattr(self._L, name)
        lock, L, errs = self._lock, self._L, self._errors

        def call(*args):
            with lock:
                rc = fn(*args)
                # a failing call on a context handle: read its message before another thread's call replaces it
                if isinstance(rc, int) and rc < 0 and args and isinstance(args[0], C.c_void_p) and args[0].value:
                    try:
                        msg = L.phy_last_error(args[0])
                        errs[threading.get_ident()] = msg.decode() if msg else ""
                    except Exception:
                        pass
                return rc
        setattr(self, name, call)
        return call

    def take_error(self):
        """The message captured with this thread's last failing call (None when there is none)."""
        import threading
        return self._errors.pop(threading.get_ident(), None)


class Matcher:
    def __init__(self, device: int = 0, hbm_budget: int = 0):
        import threading
        self._lock = threading.RLock()
        self._L = _SerializedLib(_lib.load(), self._lock)
        self._ctx = C.c_void_p()
        _lib.check(self._L.phy_ctx_create(C.byref(self._ctx), device, hbm_budget))
        self.device = device
        self.indexes: dict[int, ResidentIndex] = {}
        self.records: list = []
        self._seq_keepalive = None

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if self._ctx:
            if not getattr(self, "_release_at_exit", False):
                self._L.phy_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def release_at_exit(self):
        """One-shot command-line runs: skip the per-buffer cudaFree of close(); the driver releases
        the whole context when the process exits (freeing ~100 GB index by index costs up to a second)."""
        self._release_at_exit = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, code):
        if code != _lib.PHY_OK:
            msg = self._L.take_error()
            if msg is not None:
                raise _lib.PhylignCudaError(code, msg)
        _lib.check(code, self._ctx)

    # ------------------------------------------------------------------ index store
    def _begin(self, batch, hdr):
        idx = C.c_int(-1)
        self._ck(self._L.phy_index_begin(self._ctx, batch.encode(), hdr.term_size, hdr.canonicalize,
                                         hdr.signature_size, hdr.num_hashes, hdr.n_docs, C.byref(idx)))
        return idx.value

    def _push(self, idx_id, chunk):
        n = len(chunk)
        if n == 0:
            return
        if isinstance(chunk, (bytes, bytearray)):
            buf = (C.c_char * n).from_buffer_copy(chunk) if isinstance(chunk, bytes) else \
                (C.c_char * n).from_buffer(chunk)
        else:
            buf = (C.c_char * n).from_buffer(chunk)
        self._ck(self._L.phy_index_push(self._ctx, idx_id, buf, n))

    def load_index(self, path, batch: str | None = None) -> int:
        """Stream a `.cobs_classic[.xz]` file (or pipe) into HBM; returns the index id."""
        if batch is None:
            batch = os.path.basename(str(path)).split(".cobs_classic")[0]
        if not str(path).endswith(".xz") and os.path.isfile(path):     # plain file: the library's file loader
            return self.load_indexes([path], [batch], workers=8)[0]
        with IndexStream(path) as st:
            idx_id = self._begin(batch, st.header)
            try:
                for chunk in st.body_chunks():
                    self._push(idx_id, chunk)
                self._ck(self._L.phy_index_commit(self._ctx, idx_id))
            except Exception:
                self._L.phy_index_evict(self._ctx, idx_id)
                raise
        self.indexes[idx_id] = ResidentIndex(idx_id, batch, st.header)
        return idx_id

    def load_indexes(self, paths, batches=None, workers: int = 4, keep_paths=None, active: bool = True):
        """Load several index files concurrently: one xz decoder process + one host thread per
        stream, chunks read straight into page-locked buffers and pushed under a lock (the
        library is driven by one thread at a time).  Decoding is the bottleneck (~0.3 GB/s per
        stream), so throughput scales with `workers` up to the host's cores.
        keep_paths[i] (optional): also leave the decompressed stream there (written as .tmp, then
        renamed) -- the file the reference's `decompress_cobs` rule produces (Snakefile:364-387).
        active=False: the indexes stay out of phy_match_run until set_active_only() names them.
        Returns the index ids in the order of `paths`."""
        import threading
        from concurrent.futures import ThreadPoolExecutor
        batches = batches or [os.path.basename(str(p)).split(".cobs_classic")[0] for p in paths]
        keep_paths = keep_paths or [None] * len(paths)
        lock = threading.Lock()
        file_gate = threading.Lock()
        chunk = 8 << 20

        raw_lib = _lib.load()

        def one_file(path, batch):
            """Decompressed file: header parsed here, body by the library's file loader (reader threads ->
            page-locked ring -> DMA + re-stride on the upload stream).  The long call runs outside the
            Matcher lock: it touches only its own index entry and the upload stream."""
            st = IndexStream(path)
            hdr = st.header
            st.abort()
            with lock:
                idx_id = self._begin(batch, hdr)
                if not active:
                    self._ck(self._L.phy_index_set_active(self._ctx, idx_id, 0))
            try:
                with file_gate:          # one file at a time: the loader parallelises inside
                    _lib.check(raw_lib.phy_index_load_file(self._ctx, idx_id, os.fsencode(path), hdr.header_size,
                                                           max(1, workers)), self._ctx)
                with lock:
                    self._ck(self._L.phy_index_commit(self._ctx, idx_id))
            except Exception:
                with lock:
                    self._L.phy_index_evict(self._ctx, idx_id)
                raise
            with lock:
                self.indexes[idx_id] = ResidentIndex(idx_id, batch, hdr)
            return idx_id

        def one(path, batch, keep):
            if not str(path).endswith(".xz") and os.path.isfile(path) and not keep:
                return one_file(path, batch)
            bufs = [PinnedBuffer(chunk), PinnedBuffer(chunk)]
            tee = None
            with IndexStream(path, chunk_bytes=chunk) as st:
                with lock:
                    idx_id = self._begin(batch, st.header)
                    if not active:
                        self._ck(self._L.phy_index_set_active(self._ctx, idx_id, 0))
                try:
                    if keep:
                        os.makedirs(os.path.dirname(os.path.abspath(keep)), exist_ok=True)
                        tee = open(keep + ".tmp", "wb")
                        tee.write(st.header.to_bytes())
                    k = 0
                    for piece in st.body_chunks():
                        n = len(piece)
                        if n == 0:
                            continue
                        b = bufs[k & 1]
                        k += 1
                        b.array[:n] = np.frombuffer(piece, dtype=np.uint8)
                        if tee:
                            tee.write(b.array[:n])
                        with lock:
                            self._ck(self._L.phy_index_push(self._ctx, idx_id, b.ptr, n))
                    with lock:
                        self._ck(self._L.phy_index_commit(self._ctx, idx_id))
                    if tee:
                        tee.close()
                        tee = None
                        os.replace(keep + ".tmp", keep)
                except Exception:
                    with lock:
                        self._L.phy_index_evict(self._ctx, idx_id)
                    if tee:
                        tee.close()
                    if keep and os.path.exists(keep + ".tmp"):
                        os.unlink(keep + ".tmp")
                    raise
            with lock:
                self.indexes[idx_id] = ResidentIndex(idx_id, batch, st.header)
            return idx_id

        with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
            return list(ex.map(one, paths, batches, keep_paths))

    def set_active_only(self, idx_ids):
        """Only these resident indexes take part in the following match_run calls."""
        only = set(idx_ids)
        for i in list(self.indexes):
            self._ck(self._L.phy_index_set_active(self._ctx, i, int(i in only)))

    def budget_bytes(self) -> int:
        """HBM bytes this context may use in total (phy_ctx_create's hbm_budget, or what was free)."""
        b = C.c_uint64()
        self._ck(self._L.phy_ctx_budget(self._ctx, C.byref(b), None))
        return b.value

    def load_index_bytes(self, raw: bytes, batch: str) -> int:
        hdr, body = parse_bytes(raw)
        idx_id = self._begin(batch, hdr)
        try:
            self._push(idx_id, body)
            self._ck(self._L.phy_index_commit(self._ctx, idx_id))
        except Exception:
            self._L.phy_index_evict(self._ctx, idx_id)
            raise
        self.indexes[idx_id] = ResidentIndex(idx_id, batch, hdr)
        return idx_id

    def add_synth_index(self, batch, spec: _lib.SynthSpec, signature_size, doc_names=None,
                        num_hashes=1, canonicalize=1):
        """Bench/test utility: build an index on the device from procedural genomes."""
        from .cobs_index import ClassicHeader
        names = doc_names or [f"r{d:06d}_SYN{d:06d}" for d in range(spec.n_docs)]
        hdr = ClassicHeader(31, canonicalize, spec.n_docs, signature_size, num_hashes, names)
        idx_id = self._begin(batch, hdr)
        self._ck(self._L.phy_index_synth(self._ctx, idx_id, C.byref(spec)))
        self.indexes[idx_id] = ResidentIndex(idx_id, batch, hdr)
        return idx_id

    def insert_queries(self, idx_id, doc_of_query):
        """Test/bench utility: add the k-mers of the current queries to documents of an index
        (`cobs classic-construct` semantics); doc 0xFFFFFFFF skips a query."""
        a = np.ascontiguousarray(doc_of_query, dtype=np.uint32)
        assert len(a) == self._nq
        self._ck(self._L.phy_index_insert(self._ctx, idx_id, a.ctypes.data))

    def evict(self, idx_id):
        self._ck(self._L.phy_index_evict(self._ctx, idx_id))
        self.indexes.pop(idx_id, None)

    def index_info(self, idx_id) -> _lib.IndexInfo:
        info = _lib.IndexInfo()
        self._ck(self._L.phy_index_info_get(self._ctx, idx_id, C.byref(info)))
        return info

    def download_index(self, idx_id) -> bytes:
        n = self.indexes[idx_id].header.body_size
        buf = C.create_string_buffer(n)
        self._ck(self._L.phy_index_download(self._ctx, idx_id, buf, n))
        return buf.raw

    def download_index_into(self, idx_id, out):
        """Packed body bytes of a resident index into `out` (PinnedBuffer or uint8 numpy array of
        body_size bytes) -- no intermediate copies."""
        n = self.indexes[idx_id].header.body_size
        ptr = out.ptr if isinstance(out, PinnedBuffer) else out.ctypes.data
        assert (out.nbytes if isinstance(out, PinnedBuffer) else out.size) >= n
        self._ck(self._L.phy_index_download(self._ctx, idx_id, ptr, n))
        return n

    def write_index_file(self, idx_id, path, buf=None):
        """Write a resident index as a `.cobs_classic` file (SURVEY Appendix A.1 layout: what
        `xzcat {batch}.cobs_classic.xz` yields), tmp + rename."""
        ix = self.indexes[idx_id]
        n = ix.header.body_size
        buf = buf if buf is not None else PinnedBuffer(n)
        self.download_index_into(idx_id, buf)
        tmp = f"{path}.tmp.{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(ix.header.to_bytes())
            f.write(memoryview(buf.array if isinstance(buf, PinnedBuffer) else buf)[:n])
        os.replace(tmp, path)

    def set_ranks(self, all_batches=None):
        """Upload the integer ranks behind the merge key (-kmers, batch, ref) of
        filter_queries.py:135.  `all_batches`: every batch name of the job (all GPUs)."""
        names = sorted(set(all_batches) if all_batches is not None
                       else {ix.batch for ix in self.indexes.values()})
        rank_of = {b: i for i, b in enumerate(names)}
        for ix in self.indexes.values():
            refs = [ref_of(n) for n in ix.doc_names]
            order = sorted(range(len(refs)), key=lambda d: refs[d])
            rr = np.empty(len(refs), dtype=np.uint32)
            rr[order] = np.arange(len(refs), dtype=np.uint32)
            ix.batch_rank = rank_of[ix.batch]
            self._ck(self._L.phy_index_set_ranks(self._ctx, ix.idx_id, ix.batch_rank, rr.ctypes.data))
        self.batch_names = names
        return names

    # ------------------------------------------------------------------ queries
    def set_queries(self, records):
        """records = [(header, seq str|bytes)]; sequences must be upper-case ACGT."""
        self.records = [(h, s if isinstance(s, bytes) else s.encode()) for h, s in records]
        cat = b"".join(s for _, s in self.records)
        offs = np.zeros(len(self.records) + 1, dtype=np.uint64)
        if self.records:
            offs[1:] = np.cumsum([len(s) for _, s in self.records], dtype=np.uint64)
        self.set_queries_raw(cat, offs)

    def set_queries_raw(self, cat, offs):
        """cat: bytes / numpy uint8 array / PinnedBuffer with all bases; offs: uint64[nq+1]
        (numpy array or PinnedBuffer).  Pinned inputs are DMA'd without a staging copy."""
        def addr(x):
            if isinstance(x, PinnedBuffer):
                return x.ptr
            if isinstance(x, np.ndarray):
                return x.ctypes.data
            return C.cast(C.c_char_p(x), C.c_void_p).value
        if isinstance(offs, PinnedBuffer):
            nq = offs.nbytes // 8 - 1
        else:
            offs = np.ascontiguousarray(offs, dtype=np.uint64)
            nq = len(offs) - 1
        self._seq_keepalive = (cat, offs)
        self._nq = nq
        self._ck(self._L.phy_queries_set(self._ctx, addr(cat), addr(offs), nq))

    def fix_bases(self, bases):
        """rule fix_query's base transform on the device, in place: `bases` = writable uint8 numpy array
        (or bytearray) of sequence letters; upper-case, non-ACGT -> 'A' (Snakefile:326-332)."""
        arr = np.frombuffer(bases, dtype=np.uint8) if not isinstance(bases, np.ndarray) else bases
        if arr.size:
            self._ck(self._L.phy_fix_bases(self._ctx, arr.ctypes.data, arr.size))
        return bases

    # ------------------------------------------------------------------ match
    def match_run(self, threshold: float, top_n: int = 0, floor_mode: bool = False, merge_top_n: int = 0):
        p = _lib.MatchParams(float(threshold), int(top_n), int(bool(floor_mode)))
        self._ck(self._L.phy_match_run(self._ctx, C.byref(p), int(merge_top_n)))

    def fetch(self) -> MatchResult:
        rp = C.POINTER(_lib.Results)()
        self._ck(self._L.phy_results_fetch(self._ctx, C.byref(rp)))
        owner = _Owned(rp, self._L.phy_results_free)
        r = rp.contents
        nu, nh, nq = int(r.n_units), int(r.n_hits), int(r.n_queries)
        units = _view(r.units, nu, UNIT_DT, owner)
        hits = _view(r.hits, nh, HIT_DT, owner)
        nk = np.frombuffer(C.string_at(r.n_kmers, nq * 4), dtype=np.uint32) if nq else np.zeros(0, np.uint32)
        return MatchResult(units, hits, nk, nq, int(r.h2d_bytes), int(r.d2h_bytes), owner)

    def match(self, threshold: float, top_n: int = 0, floor_mode: bool = False, merge_top_n: int = 0,
              only=None):
        """Run + fetch.  `only`: index ids to query (the other resident indexes sit this one out)."""
        if only is None:
            self.match_run(threshold, top_n, floor_mode, merge_top_n)
            return self.fetch()
        only = set(only)
        try:
            for i in self.indexes:
                self._ck(self._L.phy_index_set_active(self._ctx, i, int(i in only)))
            self.match_run(threshold, top_n, floor_mode, merge_top_n)
            return self.fetch()
        finally:
            for i in self.indexes:
                self._L.phy_index_set_active(self._ctx, i, 1)

    def merged(self):
        """(offs[nq+1], cands structured array) of the cross-index top-N + ties merge.
        Zero-copy views of pinned library memory; the arrays (and their slices) hold a reference
        to the library block (see `_view`), so they stay valid after the next merged()/merge_host()
        call.  `self._merged_owner` is the most recent block (used by the FASTA formatter)."""
        mp = C.POINTER(_lib.Merged)()
        self._ck(self._L.phy_merged_fetch(self._ctx, C.byref(mp)))
        owner = _Owned(mp, self._L.phy_merged_free)
        m = mp.contents
        nq = int(m.n_queries)
        offs = _view(m.offs, nq + 1, np.dtype("<u8"), owner)
        cands = _view(m.cands, int(offs[-1]), CAND_DT, owner)
        self._merged_owner = owner
        return offs, cands

    def merged_range(self):
        """(q_lo, q_hi): the queries whose merged lists this rank holds (see phy_merged_range)."""
        lo, hi = C.c_uint32(), C.c_uint32()
        self._ck(self._L.phy_merged_range(self._ctx, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def merge_host(self, offs: np.ndarray, cands: np.ndarray, top_n: int):
        """filter_queries.py entry: merge host-supplied candidates (CAND_DT, grouped by query)."""
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        cands = np.ascontiguousarray(cands, dtype=CAND_DT)
        self._ck(self._L.phy_merge_host(self._ctx, len(offs) - 1, int(top_n), offs.ctypes.data,
                                        cands.ctypes.data))
        return self.merged()

    def scores(self, idx_id) -> np.ndarray:
        nq = self._nq
        d = self.indexes[idx_id].header.n_docs
        out = np.zeros((nq, d), dtype=np.uint32)
        self._ck(self._L.phy_scores(self._ctx, idx_id, out.ctypes.data))
        return out

    # ------------------------------------------------------------------ timing / bench hooks
    def timer_start(self):
        self._ck(self._L.phy_timer_start(self._ctx))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self._L.phy_timer_stop(self._ctx, C.byref(ms)))
        return ms.value

    def sync(self):
        self._ck(self._L.phy_sync(self._ctx))

    def flush_l2(self):
        self._ck(self._L.phy_flush_l2(self._ctx))

    def gathered_bytes(self) -> int:
        b = C.c_uint64()
        self._ck(self._L.phy_last_gather_bytes(self._ctx, C.byref(b)))
        return b.value

    def gathered_bytes_of(self, idx_id) -> int:
        b = C.c_uint64()
        self._ck(self._L.phy_last_gather_bytes_of(self._ctx, idx_id, C.byref(b)))
        return b.value

    def set_option(self, name: str, value: int):
        self._ck(self._L.phy_ctx_set_option(self._ctx, name.encode(), int(value)))

    def phase_ms(self):
        a = (C.c_float * 4)()
        self._ck(self._L.phy_last_phase_ms(self._ctx, a))
        return list(a)

    def synth_reads(self, specs, reads_seed, first_read, n_reads, read_len, random_q8=51, err_q16=655):
        arr = (_lib.SynthSpec * len(specs))(*specs)
        out = C.create_string_buffer(n_reads * read_len)
        self._ck(self._L.phy_synth_reads(self._ctx, arr, len(specs), reads_seed, first_read, n_reads,
                                         read_len, random_q8, err_q16, out))
        return out.raw

    def nccl_init(self, uid: bytes, rank: int, n_ranks: int):
        self._ck(self._L.phy_nccl_init(self._ctx, uid, rank, n_ranks))


def _nccl_finalize(self):
    """Collective: every rank of the job calls it at the same point (orderly communicator shutdown)."""
    self._ck(self._L.phy_nccl_finalize(self._ctx))


Matcher.nccl_finalize = _nccl_finalize


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(_lib.NCCL_ID_BYTES)
    _lib.check(_lib.load().phy_nccl_unique_id(buf))
    return buf.raw
