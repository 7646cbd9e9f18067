"""Pure-Python twins of the library's C++ text formatters / parser (test references only).

They restate the text protocols line by line the way the consumers read them
(/root/reference/scripts/postprocess_cobs.py:16-18, filter_queries.py:27-66,152-156) so that
the native code paths (phy_format_cobs_text, phy_format_filter_fasta, phy_parse_match_text,
phy_write_match_blocks) can be checked byte for byte without a GPU.  Not shipped with the product.
"""
import gzip


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
