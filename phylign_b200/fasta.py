"""Query file readers (host side).

`read_cobs_records` follows cobs's own FASTA handling in `cobs query -f` (SURVEY.md
Appendix A.8): a line starting with '>' or ';' opens a record, following lines are
concatenated, empty lines are skipped, records without sequence are dropped.
`read_fastx` is the FASTA/FASTQ reader the merge step needs (same record semantics as
readfq used by /root/reference/scripts/filter_queries.py:69-102: name = header up to the
first space).
"""
from __future__ import annotations

import gzip


def _open_text(path):
    p = str(path)
    if p.endswith(".gz"):
        return gzip.open(p, "rt")
    return open(p, "r")


def read_cobs_records(path):
    """[(header_without_first_char, seq)] in file order, empty-sequence records dropped."""
    recs = []
    head, parts = None, []
    with _open_text(path) as f:
        for line in f:
            line = line.rstrip("\r\n")
            if not line:
                continue
            if line[0] in ">;":
                if head is not None and parts:
                    recs.append((head, "".join(parts)))
                head, parts = line[1:], []
            elif head is not None:
                parts.append(line)
    if head is not None and parts:
        recs.append((head, "".join(parts)))
    return recs


def read_fastx(path):
    """[(name, seq)] for FASTA or FASTQ; name = header up to the first space."""
    recs = []
    with _open_text(path) as f:
        lines = iter(f)
        last = None
        while True:
            if last is None:
                for ln in lines:
                    if ln[0] in ">@":
                        last = ln.rstrip("\r\n")
                        break
            if last is None:
                break
            name, is_fq = last[1:].partition(" ")[0], last[0] == "@"
            last = None
            seqs = []
            for ln in lines:
                if ln[0] in "@+>":
                    last = ln.rstrip("\r\n")
                    break
                seqs.append(ln.rstrip("\r\n"))
            seq = "".join(seqs)
            recs.append((name, seq))
            if last is not None and last[0] == "+":     # FASTQ: skip the quality block
                got = 0
                last = None
                for ln in lines:
                    got += len(ln.rstrip("\r\n"))
                    if got >= len(seq):
                        break
            elif last is None:
                break
    return recs


# ---- query preparation ---------------------------------------------------------
# What /root/reference/Snakefile:326-332 does with `seqtk seq -A -U -C | awk gsub(/[^ACGT]/,"A")`:
# FASTA out, one line per sequence, upper case, comments dropped, every other letter -> A.
_FIX = bytes((c if chr(c) in "ACGT" else ord("A")) for c in
             (ord(chr(b).upper()) if b < 128 else b for b in range(256)))


def fix_query_seq(seq) -> str:
    """Upper-case and replace everything outside ACGT by 'A' (table driven, C speed)."""
    b = seq.encode() if isinstance(seq, str) else bytes(seq)
    return b.translate(_FIX).decode()


def fix_query_file(path) -> str:
    """The content of intermediate/00_queries_preprocessed/{qfile}.fa for one input file."""
    return "".join(f">{name}\n{fix_query_seq(seq)}\n" for name, seq in read_fastx(path))
