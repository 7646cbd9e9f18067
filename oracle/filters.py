"""CPU ORACLE for the two Python filters of the match stage -- test infrastructure.

Restates, on parsed data instead of text streams:
  * ``/root/reference/scripts/postprocess_cobs.py:21-38`` (per-batch top-N + ties,
    random-prefix stripping ``:16-18``),
  * ``/root/reference/scripts/filter_queries.py:27-66`` (match-file parser) and
    ``:105-156`` (``SingleQuery`` running top-N + ties, key ``(-kmers, batch, ref)``),
    ``:152-156`` (FASTA record with the candidate list).
Pinned by ``tests/golden/`` vectors that were produced by executing the
unmodified reference scripts (``tests/golden/make_golden.py``).
"""
from __future__ import annotations


def remove_rnd_id(name: str) -> str:
    """postprocess_cobs.py:16-18 -- keep "_" + text after the first underscore."""
    _, _, r = name.partition("_")
    return "_" + r


def postprocess_block(hits, keep: int):
    """postprocess_cobs.py:21-38 for one query block.

    ``hits`` = [(doc_name, score)] in cobs order (score descending).  Returns the
    printed [(stripped_name, score)].
    """
    out = []
    i = 0
    min_kmers = 0
    for name, score in hits:
        y = remove_rnd_id(name)
        i += 1
        if i < keep:
            out.append((y, score))
        elif i == keep:
            out.append((y, score))
            min_kmers = score
        elif score == min_kmers:
            out.append((y, score))
    return out


def parse_cobs_text(text: str):
    """Split `cobs query` / match-file text into [(header, n, [(name, score)])]."""
    blocks = []
    for line in text.splitlines():
        if not line:
            continue
        if line[0] == "*":
            head, _, n = line[1:].rpartition("\t")
            blocks.append((head, int(n), []))
        else:
            name, _, score = line.rpartition("\t")
            blocks[-1][2].append((name, int(score)))
    return blocks


def postprocess_text(text: str, keep: int) -> str:
    """Whole-stream equivalent of ``postprocess_cobs.py -n keep``."""
    out = []
    for head, n, hits in parse_cobs_text(text):
        out.append(f"*{head}\t{n}\n")
        out.extend(f"{nm}\t{sc}\n" for nm, sc in postprocess_block(hits, keep))
    return "".join(out)


class SingleQuery:
    """filter_queries.py:105-156, restated line by line."""

    def __init__(self, qname, seq, keep):
        self.keep = keep
        self.min_kmers = 0
        self.matches = []
        self.qname = qname
        self.seq = seq

    def add_matches(self, batch, matches):
        for ref, kmers in matches:
            kmers = int(kmers)
            if kmers >= self.min_kmers:
                self.matches.append((batch, ref, kmers))
        self.matches.sort(key=lambda x: (-x[2], x[0], x[1]))
        losers = self.matches[self.keep:]
        self.matches = self.matches[:self.keep]
        if losers:
            self.min_kmers = self.matches[-1][2]
            for x in losers:
                if x[2] == self.min_kmers:
                    self.matches.append(x)
                else:
                    break

    def record(self):
        return f">{self.qname} {','.join(x[1] for x in self.matches)}\n{self.seq}"


def merge_running(queries, batches, keep: int) -> str:
    """filter_queries.py process_files: ``queries`` = [(qname, seq)] in file order,
    ``batches`` = [(batch_name, [(qname, [(ref, kmers)])])] in argv order."""
    d = {}
    for qname, seq in queries:
        d[qname] = SingleQuery(qname, seq, keep)
    for batch, per_query in batches:
        for qname, matches in per_query:
            d[qname].add_matches(batch, matches)
    return "".join(q.record() + "\n" for q in d.values())


def merge_closed_form(queries, batches, keep: int) -> str:
    """Closed form of the running merge: all candidates whose score
    is >= the keep-th largest score over all batches, ordered by
    (score desc, batch asc, ref asc).  Tests prove it equals ``merge_running``."""
    cand = {q: [] for q, _ in queries}
    for batch, per_query in batches:
        for qname, matches in per_query:
            cand[qname].extend((batch, ref, int(k)) for ref, k in matches)
    out = []
    for qname, seq in queries:
        c = sorted(cand[qname], key=lambda x: (-x[2], x[0], x[1]))
        if len(c) > keep:
            cut = c[keep - 1][2]
            c = [x for x in c if x[2] >= cut]
        out.append(f">{qname} {','.join(x[1] for x in c)}\n{seq}\n")
    return "".join(out)
