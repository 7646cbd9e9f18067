"""COBS classic index container: header parsing and body streaming (host side).

Layout (the file `cobs query -i` reads at
/root/reference/scripts/run_cobs_streaming.sh:24-29):
    "COBS:" "CLASSIC_INDEX" u32 version=1 | u32 term_size | u8 canonicalize | u32 n_docs |
    u64 signature_size | u64 num_hashes | n_docs x (name "\\n") | "CLASSIC_INDEX" | body
body = signature_size rows of ceil(n_docs/8) bytes; doc d <-> byte d/8, bit d%8 of its row.
The index may arrive through a pipe (`-i <(xzcat ...)`), so parsing is strictly sequential.
"""
from __future__ import annotations

import io
import lzma
import os
import shutil
import struct
import subprocess
from dataclasses import dataclass, field

MAGIC0 = b"COBS:"
MAGIC1 = b"CLASSIC_INDEX"


class IndexFormatError(ValueError):
    pass


@dataclass
class ClassicHeader:
    term_size: int
    canonicalize: int
    n_docs: int
    signature_size: int
    num_hashes: int
    doc_names: list = field(default_factory=list)
    header_size: int = 0

    @property
    def row_size(self) -> int:
        return (self.n_docs + 7) // 8

    @property
    def body_size(self) -> int:
        return self.signature_size * self.row_size

    def to_bytes(self) -> bytes:
        out = [MAGIC0, MAGIC1, struct.pack("<IIBIQQ", 1, self.term_size, self.canonicalize,
                                           self.n_docs, self.signature_size, self.num_hashes)]
        out += [n.encode() + b"\n" for n in self.doc_names]
        out.append(MAGIC1)
        return b"".join(out)


def _read_exact(f, n: int) -> bytes:
    chunks = []
    while n:
        b = f.read(n)
        if not b:
            raise IndexFormatError("truncated COBS classic index header")
        chunks.append(b)
        n -= len(b)
    return b"".join(chunks)


def read_header(f) -> ClassicHeader:
    """Parse the header from a binary stream positioned at byte 0."""
    if _read_exact(f, 5) != MAGIC0 or _read_exact(f, 13) != MAGIC1:
        raise IndexFormatError("not a COBS classic index (magic mismatch)")
    version, k, canon, n_docs, sig, nh = struct.unpack("<IIBIQQ", _read_exact(f, 29))
    if version != 1:
        raise IndexFormatError(f"unsupported classic index version {version}")
    names = []
    size = 5 + 13 + 29
    buf, pos = b"", 0               # names are cut out of `buf` at a moving offset (no re-copying)
    while len(names) < n_docs:
        nl = buf.find(b"\n", pos)
        if nl < 0:
            more = f.read(65536)   # may run into the body: the excess is carried over
            if not more:
                raise IndexFormatError("truncated document name table")
            buf = buf[pos:] + more
            pos = 0
            continue
        names.append(buf[pos:nl].decode())
        size += nl + 1 - pos
        pos = nl + 1
    buf = buf[pos:]
    # `buf` may already hold bytes past the name table: end magic (+ body)
    need = 13 - len(buf)
    if need > 0:
        buf += _read_exact(f, need)
    if buf[:13] != MAGIC1:
        raise IndexFormatError("end-of-header magic missing")
    hdr = ClassicHeader(k, canon, n_docs, sig, nh, names, size + 13)
    hdr._carry = buf[13:]          # body bytes read ahead
    return hdr


class IndexStream:
    """Sequential reader of one classic index: header, then body chunks."""

    def __init__(self, path: str, chunk_bytes: int = 16 << 20):
        self.path = path
        self.chunk_bytes = chunk_bytes
        self._proc = None
        p = os.fspath(path)
        if p.endswith(".xz"):
            xz = shutil.which("xzcat")
            if xz:  # ~1.5x faster than python's lzma, and decodes in its own process
                # same flags as the reference (run_cobs_streaming.sh:27)
                self._proc = subprocess.Popen([xz, "--no-sparse", "--ignore-check", p],
                                              stdout=subprocess.PIPE, bufsize=0)
                self._f = self._proc.stdout
            else:
                self._f = lzma.open(p, "rb")
        else:
            self._f = open(p, "rb", buffering=0)
        self.header = read_header(self._f)
        self._carry = self.header._carry
        self._sent = 0

    def body_chunks(self):
        """Yield the body as bytes-like chunks; checks the size invariant of A.1."""
        total = self.header.body_size
        if self._carry:
            c = self._carry[:total]
            self._sent += len(c)
            extra = len(self._carry) - len(c)
            self._carry = b""
            yield c
            if extra:
                raise IndexFormatError("index holds bytes past signature_size*row_size")
        buf = bytearray(self.chunk_bytes)
        view = memoryview(buf)
        while self._sent < total:
            want = min(self.chunk_bytes, total - self._sent)
            n = self._f.readinto(view[:want])
            if not n:
                raise IndexFormatError(
                    f"index body truncated: {self._sent} of {total} bytes "
                    f"(signature_size*row_size)")
            self._sent += n
            yield view[:n]
        if self._f.read(1):
            raise IndexFormatError("index holds bytes past signature_size*row_size")

    def close(self):
        try:
            self._f.close()
        finally:
            if self._proc is not None:
                rc = self._proc.wait()
                self._proc = None
                if rc != 0:
                    raise IndexFormatError(f"xzcat failed on {self.path} (exit {rc})")

    def abort(self):
        """Stop reading (header-only use): close the stream and reap the decoder."""
        try:
            self._f.close()
            if self._proc is not None:
                self._proc.kill()
                self._proc.wait()
                self._proc = None
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.close()
        else:
            try:
                self._f.close()
                if self._proc is not None:
                    self._proc.kill()
                    self._proc.wait()
            except Exception:
                pass


def parse_bytes(raw: bytes):
    """(header, body bytes) of an in-memory classic index; validates the size invariant."""
    f = io.BytesIO(raw)
    hdr = read_header(f)
    body = hdr._carry + f.read()
    if len(body) != hdr.body_size:
        raise IndexFormatError(f"index body is {len(body)} bytes, expected {hdr.body_size}")
    return hdr, body


def ref_of(doc_name: str) -> str:
    """Accession of a Phylign doc name "<rnd>_<acc>" (postprocess_cobs.py:16-18)."""
    return doc_name.partition("_")[2]
