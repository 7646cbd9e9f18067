"""Text protocols of the match stage (host side).

* `format_cobs_text`  -> what `cobs query` prints: per query
  "*<header>\\t<n_pass>" then "<doc_name>\\t<score>" lines, score descending.
  With strip_prefix=True the doc names are written the way
  /root/reference/scripts/postprocess_cobs.py:16-18 rewrites them ("_" + accession), i.e.
  the content of intermediate/03_match/{batch}____{qfile}.gz.
* `format_filter_fasta` -> what /root/reference/scripts/filter_queries.py:152-156,195-199
  prints: ">{qname} {ref1,ref2,...}\\n{seq}".
"""
from __future__ import annotations

import numpy as np


def format_cobs_text(records, result, index, strip_prefix: bool = False) -> str:
    """records = [(header, seq)] as given to set_queries; index = ResidentIndex."""
    names = index.doc_names
    if strip_prefix:
        names = ["_" + n.partition("_")[2] for n in names]
    units = result.units_of(index.idx_id)
    by_query = {int(u["query"]): u for u in units}
    hits = result.hits
    out = []
    for q, (head, seq) in enumerate(records):
        if len(seq) == 0:      # cobs never runs a record without sequence (A.8)
            continue
        u = by_query.get(q)
        if u is None:
            out.append(f"*{head}\t0\n")
            continue
        out.append(f"*{head}\t{int(u['n_pass'])}\n")
        o, n = int(u["offset"]), int(u["n_kept"])
        h = hits[o:o + n]
        out.extend(f"{names[d]}\t{s}\n" for d, s in zip(h["doc"].tolist(), h["score"].tolist()))
    return "".join(out)


def format_filter_fasta(records, offs, cands, ref_names_by_rank) -> str:
    """records = [(qname, seq str)]; ref_names_by_rank[batch_rank][doc] = accession."""
    out = []
    br = cands["batch_rank"].tolist()
    dc = cands["doc"].tolist()
    o = offs.tolist()
    for q, (qname, seq) in enumerate(records):
        refs = [ref_names_by_rank[br[i]][dc[i]] for i in range(o[q], o[q + 1])]
        out.append(f">{qname} {','.join(refs)}\n{seq}\n")
    return "".join(out)
