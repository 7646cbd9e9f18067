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
import os


def _open_text(path):
    # newline="\n": lines end at LF only, like std::getline in cobs and like phy_fasta_read (a lone CR
    # inside a line is data, a trailing CR is stripped by the readers)
    p = str(path)
    if p.endswith(".gz"):
        return gzip.open(p, "rt", newline="\n")
    return open(p, "r", newline="\n")


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


# ---- flat-array query file (native reader) -----------------------------------------------------------
class QueryFile:
    """A query FASTA as four flat arrays (phy_fasta_read, include/phylign_cuda.h): `seqs`/`soffs`
    (concatenated bases, page-locked when a GPU is present) and `headers`/`hoffs` (header lines
    without '>'), plus `name_len` (query name = header up to the first blank).  Same record rules as
    read_cobs_records.  Gzipped input goes through the Python reader and is packed the same way."""

    def __init__(self, path):
        import ctypes as C
        import numpy as np
        from . import _lib
        self._fp = None
        self.path = str(path)
        if self.path.endswith(".gz"):
            recs = read_cobs_records(self.path)
            hs = [h.encode() for h, _ in recs]
            ss = [s.encode() for _, s in recs]
            self.n = len(recs)
            self.seqs = np.frombuffer(b"".join(ss) + b"\0", dtype=np.uint8)
            self.headers = np.frombuffer(b"".join(hs) + b"\0", dtype=np.uint8)
            self.soffs = np.zeros(self.n + 1, np.uint64)
            self.hoffs = np.zeros(self.n + 1, np.uint64)
            if recs:
                self.soffs[1:] = np.cumsum([len(s) for s in ss], dtype=np.uint64)
                self.hoffs[1:] = np.cumsum([len(h) for h in hs], dtype=np.uint64)
            self.name_len = np.array([len(h.split(b" ")[0]) for h in hs], dtype=np.uint32)
            self.simple = False
            return
        L = _lib.load()
        fp = C.POINTER(_lib.Fasta)()
        rc = L.phy_fasta_read(os.fsencode(self.path), C.byref(fp))
        if rc != 0:
            raise OSError(L.phy_last_error(None).decode())
        self._fp, self._L = fp, L
        f = fp.contents
        self.n = int(f.n)
        self.simple = bool(f.simple)

        def view(addr, nbytes, dt):
            if nbytes == 0:
                return np.zeros(0, dt)
            buf = (C.c_char * nbytes).from_address(addr)
            buf._phy_owner = self          # the arrays keep the native block alive
            return np.frombuffer(buf, dtype=dt)
        self.soffs = view(C.addressof(f.soffs.contents), (self.n + 1) * 8, np.uint64)
        self.hoffs = view(C.addressof(f.hoffs.contents), (self.n + 1) * 8, np.uint64)
        self.name_len = view(C.addressof(f.name_len.contents), self.n * 4, np.uint32) if self.n else np.zeros(0, np.uint32)
        self.seqs = view(f.seqs, int(self.soffs[-1]) + 1, np.uint8)
        self.headers = view(f.headers, int(self.hoffs[-1]) + 1, np.uint8)

    def __del__(self):
        try:
            if self._fp:
                self._L.phy_fasta_free(self._fp)
                self._fp = None
        except Exception:
            pass

    @property
    def total_bases(self) -> int:
        return int(self.soffs[-1])

    def header(self, q) -> str:
        return bytes(self.headers[int(self.hoffs[q]):int(self.hoffs[q + 1])]).decode()

    def names(self):
        hb = bytes(self.headers)
        ho, nl = self.hoffs.tolist(), self.name_len.tolist()
        return [hb[ho[q]:ho[q] + nl[q]].decode() for q in range(self.n)]

    def records(self):
        """[(header, seq bytes)] -- the list form the older entry points take."""
        hb, sb = bytes(self.headers), bytes(self.seqs)
        ho, so = self.hoffs.tolist(), self.soffs.tolist()
        return [(hb[ho[q]:ho[q + 1]].decode(), sb[so[q]:so[q + 1]]) for q in range(self.n)]

    def block_ranges(self, max_bases: int):
        """Consecutive [q0, q1) ranges of at most max_bases bases (at least one record each)."""
        import numpy as np
        out, q0 = [], 0
        so = self.soffs
        while q0 < self.n:
            q1 = int(np.searchsorted(so, so[q0] + np.uint64(max_bases), "right")) - 1
            q1 = min(self.n, max(q1, q0 + 1))
            out.append((q0, q1))
            q0 = q1
        return out or [(0, 0)]
