"""Text protocols of the match stage (host side).

* `format_cobs_text_fast`  -> what `cobs query` prints: per query
  "*<header>\\t<n_pass>" then "<doc_name>\\t<score>" lines, score descending.
  With strip_prefix=True the doc names are written the way
  /root/reference/scripts/postprocess_cobs.py:16-18 rewrites them ("_" + accession), i.e.
  the content of intermediate/03_match/{batch}____{qfile}.gz.
* `format_filter_fasta_fast` -> what /root/reference/scripts/filter_queries.py:152-156,195-199
  prints: ">{qname} {ref1,ref2,...}\\n{seq}".
"""
from __future__ import annotations

import numpy as np


# ---- fast paths: the same two formats produced by the library's C++ formatters -----------------
def _cat(strings):
    """(concatenated bytes, uint64 offsets[n+1]) of a list of str/bytes."""
    bs = [x if isinstance(x, bytes) else x.encode() for x in strings]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    return b"".join(bs), offs


_HDR_CACHE = (None, None, None)


def format_cobs_text_fast(records, result, index, strip_prefix: bool = False, results_ptr=None) -> bytes:
    """The cobs text of one index for `records` = [(header, seq)], via phy_format_cobs_text.
    `results_ptr`: ctypes POINTER(Results) (defaults to the one backing `result`)."""
    import ctypes as C
    from . import _lib
    L = _lib.load()
    rp = results_ptr if results_ptr is not None else result._owner.ptr
    global _HDR_CACHE          # the same record list is formatted once per index: build its arrays once
    if _HDR_CACHE[0] is not records:
        _HDR_CACHE = (records, _cat([h for h, _ in records]),
                      np.array([len(s) == 0 for _, s in records], dtype=np.uint8))
    (hcat, hoffs), skip = _HDR_CACHE[1], _HDR_CACHE[2]
    if not hasattr(index, "_names_cat"):
        index._names_cat = _cat(index.doc_names)
    ncat, noffs = index._names_cat
    out, n = C.c_void_p(), C.c_uint64()
    _lib.check(L.phy_format_cobs_text(rp, index.idx_id, C.cast(C.c_char_p(hcat), C.c_void_p), hoffs.ctypes.data,
                                      skip.ctypes.data, ncat,
                                      noffs.ctypes.data, len(index.doc_names), int(strip_prefix),
                                      C.byref(out), C.byref(n)))
    try:
        return C.string_at(out, n.value)
    finally:
        L.phy_text_free(out)


def format_cobs_text_arrays(headers, hoffs, result, index, strip_prefix: bool = False) -> bytes:
    """The cobs text of one index for a block of a fasta.QueryFile: `headers` = its uint8 header array,
    `hoffs` = the block's slice of header offsets (nq+1).  No per-record Python objects."""
    import ctypes as C
    from . import _lib
    L = _lib.load()
    if not hasattr(index, "_names_cat"):
        index._names_cat = _cat(index.doc_names)
    ncat, noffs = index._names_cat
    hoffs = np.ascontiguousarray(hoffs, dtype=np.uint64)
    out, n = C.c_void_p(), C.c_uint64()
    _lib.check(L.phy_format_cobs_text(result._owner.ptr, index.idx_id, headers.ctypes.data, hoffs.ctypes.data, None, ncat,
                                      noffs.ctypes.data, len(index.doc_names), int(strip_prefix),
                                      C.byref(out), C.byref(n)))
    try:
        return C.string_at(out, n.value)
    finally:
        L.phy_text_free(out)


def format_filter_fasta_fast(records, merged_ptr, ref_names_by_rank) -> bytes:
    """The 04_filter FASTA for `records` = [(qname, seq)], via phy_format_filter_fasta.
    `merged_ptr`: ctypes POINTER(Merged); ref_names_by_rank[batch_rank][doc] = accession."""
    import ctypes as C
    from . import _lib
    L = _lib.load()
    qcat, qoffs = _cat([q for q, _ in records])
    scat, soffs = _cat([s for _, s in records])
    nb = (max(ref_names_by_rank) + 1) if ref_names_by_rank else 0
    cats = [_cat(ref_names_by_rank.get(b, [])) for b in range(nb)]
    name_ptrs = (C.c_char_p * max(nb, 1))(*[c for c, _ in cats])
    off_ptrs = (C.c_void_p * max(nb, 1))(*[o.ctypes.data for _, o in cats])
    counts = np.array([len(ref_names_by_rank.get(b, [])) for b in range(nb)] or [0], dtype=np.uint32)
    out, n = C.c_void_p(), C.c_uint64()
    _lib.check(L.phy_format_filter_fasta(merged_ptr, qcat, qoffs.ctypes.data, scat, soffs.ctypes.data, nb,
                                         name_ptrs, off_ptrs, counts.ctypes.data, C.byref(out), C.byref(n)))
    try:
        return C.string_at(out, n.value)
    finally:
        L.phy_text_free(out)


def _ref_tables(ref_names_by_rank):
    import ctypes as C
    nb = (max(ref_names_by_rank) + 1) if ref_names_by_rank else 0
    cats = [_cat(ref_names_by_rank.get(b, [])) for b in range(nb)]
    name_ptrs = (C.c_char_p * max(nb, 1))(*[c for c, _ in cats])
    off_ptrs = (C.c_void_p * max(nb, 1))(*[o.ctypes.data for _, o in cats])
    counts = np.array([len(ref_names_by_rank.get(b, [])) for b in range(nb)] or [0], dtype=np.uint32)
    return nb, cats, name_ptrs, off_ptrs, counts


def write_filter_fasta_native(path, merged_ptr, qf, ref_names_by_rank, q_begin: int = 0, q_end: int = 0xFFFFFFFF,
                              q_base: int = 0, append: bool = False) -> int:
    """intermediate/04_filter/{qfile}.fa written by the library (phy_write_filter_fasta) from the flat
    arrays of a fasta.QueryFile: no Python strings, tmp + rename.  [q_begin, q_end): the queries of the
    merged lists to write (a part of the file when they are sharded over GPUs); q_base: query number of
    the merged lists' first query in the file (a query block that starts at record q_base).
    append=True: `path` is the caller's own temporary file, the text is appended and nothing is renamed.
    Returns the bytes written."""
    import ctypes as C
    import os
    from . import _lib
    L = _lib.load()
    nb, cats, name_ptrs, off_ptrs, counts = _ref_tables(ref_names_by_rank)
    n = C.c_uint64()
    _lib.check(L.phy_write_filter_fasta(os.fsencode(path), merged_ptr, qf.headers.ctypes.data,
                                        qf.hoffs.ctypes.data + 8 * q_base, qf.name_len.ctypes.data + 4 * q_base,
                                        qf.seqs.ctypes.data, qf.soffs.ctypes.data + 8 * q_base, nb,
                                        name_ptrs, off_ptrs, counts.ctypes.data, int(q_begin), int(q_end), int(append),
                                        C.byref(n)))
    return n.value


# ---- 04 -> 05 hand-off -------------------------------------------------------
def candidate_buckets(qnames, offs, cands, ref_names_by_rank):
    """{batch_rank: [(ref, [qname, ...])]}: for every reference that is a candidate of at least one
    query, the queries to align against it, in query-file order.  The same mapping
    /root/reference/scripts/batch_align.py:126-171 (`load_qdicts`, rname_to_qnames) rebuilds for
    every batch by re-parsing the whole 04_filter FASTA; emitted once here."""
    n = len(cands)
    if n == 0:
        return {}
    q_of = np.repeat(np.arange(len(qnames), dtype=np.int64), np.diff(np.asarray(offs).astype(np.int64)))
    br = cands["batch_rank"].astype(np.int64)
    dc = cands["doc"].astype(np.int64)
    order = np.lexsort((q_of, dc, br))          # by batch, then reference, queries in file order
    out = {}
    i = 0
    br_s, dc_s, q_s = br[order], dc[order], q_of[order]
    cuts = np.flatnonzero((np.diff(br_s) != 0) | (np.diff(dc_s) != 0)) + 1
    for lo, hi in zip(np.concatenate(([0], cuts)), np.concatenate((cuts, [n]))):
        b, d = int(br_s[lo]), int(dc_s[lo])
        out.setdefault(b, []).append((ref_names_by_rank[b][d], [qnames[q] for q in q_s[lo:hi].tolist()]))
    return out


def format_bucket_tsv(bucket) -> str:
    """One line per reference: "<ref>\t<qname1>,<qname2>,..."."""
    return "".join(f"{ref}\t{','.join(qs)}\n" for ref, qs in bucket)
