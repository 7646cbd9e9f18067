"""CPU ORACLE bindings -- test infrastructure, NOT product code.

ctypes wrapper over ``oracle/_build/liboracle.so`` (built from ``cobs_oracle.c``
by ``oracle/Makefile``).  **parity unpinned against the real `cobs` binary**
(COBS 0.2.1 is an un-vendored conda dependency of the reference,
``/root/reference/envs/cobs.yaml:5``); see ``cobs_oracle.h`` for what is pinned.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  ``phylign_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
CLI_PATH = os.path.join(_HERE, "_build", "cobs_oracle")
NATIVE_LIB_PATH = os.path.join(_HERE, "_build", "native", "liboracle.so")
NATIVE_CLI_PATH = os.path.join(_HERE, "_build", "native", "cobs_oracle")
_use_native = False


def _stale(*outs) -> bool:
    srcs = [os.path.join(_HERE, f) for f in ("cobs_oracle.c", "cobs_oracle.h", "cobs_oracle_cli.c", "Makefile")]
    if not all(os.path.exists(o) for o in outs):
        return True
    t = min(os.path.getmtime(o) for o in outs)
    return any(os.path.getmtime(x) > t for x in srcs)


def build(force: bool = False) -> None:
    """Compile the oracle with gcc (seconds): the portable build the tests use."""
    if force or _stale(LIB_PATH, CLI_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)


def _cpu_id() -> str:
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        keep = [l for l in txt.splitlines() if l.startswith(("model name", "flags"))][:2]
        return hashlib.sha1("\n".join(keep).encode()).hexdigest()
    except Exception:
        return "unknown"


def build_native(force: bool = False) -> bool:
    """-O3 -march=native build for the CPU timing legs of bench.py, compiled on the machine that
    runs it (never shipped between machines).  Must be called before the first lib() use.
    Returns False (and stays on the portable build) when gcc is missing."""
    global _use_native
    try:
        stamp = os.path.join(_HERE, "_build", "native", ".cpu")
        cpu = _cpu_id()
        same_cpu = os.path.exists(stamp) and open(stamp).read() == cpu     # -march=native is per machine
        if force or not same_cpu or _stale(NATIVE_LIB_PATH, NATIVE_CLI_PATH):
            subprocess.check_call(["make", "-C", _HERE, "-B", "native"], stdout=subprocess.DEVNULL)
            with open(stamp, "w") as f:
                f.write(cpu)
        _use_native = _lib is None
    except Exception:
        _use_native = False
    return _use_native


def cli_path() -> str:
    return NATIVE_CLI_PATH if _use_native else CLI_PATH


class _Index(C.Structure):
    _fields_ = [("term_size", C.c_uint32), ("canonicalize", C.c_uint8),
                ("n_docs", C.c_uint32), ("signature_size", C.c_uint64),
                ("num_hashes", C.c_uint64), ("row_size", C.c_uint64),
                ("doc_names", C.POINTER(C.c_char_p)), ("body", C.POINTER(C.c_uint8))]


class SynthSpec(C.Structure):
    """Synthetic workload spec v1 (mirrors ``orc_synth``)."""
    _fields_ = [("seed", C.c_uint64), ("n_docs", C.c_uint32), ("genome_len", C.c_uint32),
                ("clade_size", C.c_uint32), ("clade_sub_q16", C.c_uint32),
                ("doc_sub_q16", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not _use_native:
            build()
        L = C.CDLL(NATIVE_LIB_PATH if _use_native else LIB_PATH)
        L.orc_simd_mode.restype = C.c_int
        L.orc_set_simd_mode.restype = C.c_int
        L.orc_set_simd_mode.argtypes = [C.c_int]
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64]
        L.orc_canonical.restype = C.c_int
        L.orc_canonical.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
        L.orc_signature_size.restype = C.c_uint64
        L.orc_signature_size.argtypes = [C.c_uint64, C.c_uint64, C.c_double]
        L.orc_index_new.restype = C.POINTER(_Index)
        L.orc_index_new.argtypes = [C.c_uint32, C.c_uint8, C.c_uint32, C.c_uint64, C.c_uint64,
                                    C.POINTER(C.c_char_p)]
        L.orc_index_free.argtypes = [C.POINTER(_Index)]
        L.orc_index_add_doc.restype = C.c_int
        L.orc_index_add_doc.argtypes = [C.POINTER(_Index), C.c_uint32, C.c_char_p, C.c_uint64]
        L.orc_index_header_size.restype = C.c_uint64
        L.orc_index_header_size.argtypes = [C.POINTER(_Index)]
        L.orc_index_write.restype = C.c_int
        L.orc_index_write.argtypes = [C.POINTER(_Index), C.c_char_p]
        L.orc_index_read.restype = C.POINTER(_Index)
        L.orc_index_read.argtypes = [C.c_char_p]
        L.orc_index_parse.restype = C.POINTER(_Index)
        L.orc_index_parse.argtypes = [C.c_char_p, C.c_uint64]
        L.orc_threshold_terms.restype = C.c_uint32
        L.orc_threshold_terms.argtypes = [C.c_double, C.c_uint32, C.c_int]
        L.orc_query_scores.restype = C.c_int64
        L.orc_query_scores.argtypes = [C.POINTER(_Index), C.c_char_p, C.c_uint64, C.c_void_p]
        L.orc_query_scores_sliced.restype = C.c_int64
        L.orc_query_scores_sliced.argtypes = [C.POINTER(_Index), C.c_char_p, C.c_uint64,
                                              C.c_void_p, C.c_int]
        L.orc_query_batch.restype = C.c_int64
        L.orc_query_batch.argtypes = [C.POINTER(_Index), C.c_char_p, C.c_void_p, C.c_uint32,
                                      C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_mix64.restype = C.c_uint64
        L.orc_mix64.argtypes = [C.c_uint64]
        L.orc_synth_base.restype = C.c_uint32
        L.orc_synth_base.argtypes = [C.POINTER(SynthSpec), C.c_uint32, C.c_uint32]
        L.orc_synth_genome.argtypes = [C.POINTER(SynthSpec), C.c_uint32, C.c_char_p]
        L.orc_synth_read.argtypes = [C.POINTER(SynthSpec), C.c_uint32, C.c_uint64, C.c_uint64,
                                     C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p]
        _lib = L
    return _lib


def set_simd(avx2: bool) -> str:
    """Counting kernel of the sliced path: "sse2" (cobs 0.2.1 shape) or "avx2"; returns what is in effect."""
    return "avx2" if lib().orc_set_simd_mode(int(bool(avx2))) else "sse2"


def xxh64(data: bytes, seed: int = 0) -> int:
    return lib().orc_xxh64(data, len(data), seed)


def canonical(kmer: bytes) -> bytes | None:
    out = C.create_string_buffer(len(kmer))
    if lib().orc_canonical(kmer, len(kmer), out) != 0:
        return None
    return out.raw


def threshold_terms(threshold: float, num_kmers: int, floor_mode: bool = False) -> int:
    return lib().orc_threshold_terms(threshold, num_kmers, int(floor_mode))


def signature_size(max_doc_kmers: int, num_hashes: int = 1, fpr: float = 0.3) -> int:
    return lib().orc_signature_size(max_doc_kmers, num_hashes, fpr)


class OracleIndex:
    """A COBS classic index held in host RAM by the oracle."""

    def __init__(self, ptr):
        if not ptr:
            raise ValueError("oracle: could not create / parse index")
        self._p = ptr

    # -- construction -----------------------------------------------------
    @classmethod
    def new(cls, n_docs, signature_size, doc_names=None, term_size=31, canonicalize=1,
            num_hashes=1):
        arr = None
        if doc_names is not None:
            arr = (C.c_char_p * n_docs)(*[n.encode() if isinstance(n, str) else n
                                          for n in doc_names])
        return cls(lib().orc_index_new(term_size, canonicalize, n_docs, signature_size,
                                       num_hashes, arr))

    @classmethod
    def construct(cls, docs, doc_names=None, term_size=31, canonicalize=1, num_hashes=1,
                  fpr=0.3, signature_size_override=None):
        """classic-construct restatement: ``docs`` = list of ASCII byte strings."""
        max_kmers = max([max(len(d) - term_size + 1, 0) for d in docs] + [1])
        sig = signature_size_override or signature_size(max_kmers, num_hashes, fpr)
        idx = cls.new(len(docs), sig, doc_names, term_size, canonicalize, num_hashes)
        for d, seq in enumerate(docs):
            idx.add_doc(d, seq)
        return idx

    @classmethod
    def read(cls, path):
        return cls(lib().orc_index_read(os.fsencode(path)))

    @classmethod
    def parse(cls, buf: bytes):
        return cls(lib().orc_index_parse(buf, len(buf)))

    def add_doc(self, d, seq: bytes):
        if lib().orc_index_add_doc(self._p, d, seq, len(seq)) != 0:
            raise ValueError("oracle: add_doc failed")

    def write(self, path):
        if lib().orc_index_write(self._p, os.fsencode(path)) != 0:
            raise OSError("oracle: cannot write " + str(path))

    def __del__(self):
        try:
            lib().orc_index_free(self._p)
        except Exception:
            pass

    # -- properties ---------------------------------------------------------
    term_size = property(lambda s: s._p.contents.term_size)
    canonicalize = property(lambda s: s._p.contents.canonicalize)
    n_docs = property(lambda s: s._p.contents.n_docs)
    signature_size = property(lambda s: s._p.contents.signature_size)
    num_hashes = property(lambda s: s._p.contents.num_hashes)
    row_size = property(lambda s: s._p.contents.row_size)
    header_size = property(lambda s: lib().orc_index_header_size(s._p))

    @property
    def doc_names(self):
        c = self._p.contents
        return [c.doc_names[d].decode() for d in range(c.n_docs)]

    @property
    def body(self) -> np.ndarray:
        """(signature_size, row_size) uint8 view of the packed rows."""
        c = self._p.contents
        n = c.signature_size * c.row_size
        a = np.ctypeslib.as_array(c.body, shape=(n,))
        return a.reshape(c.signature_size, c.row_size)

    # -- query ---------------------------------------------------------------
    def scores(self, seq: bytes, sliced: bool = False, threads: int = 1):
        """(K, scores[D]); K = -1 when the query holds a non-ACGT letter."""
        out = np.zeros(self.n_docs, dtype=np.uint32)
        if sliced:
            k = lib().orc_query_scores_sliced(self._p, seq, len(seq), out.ctypes.data, threads)
        else:
            k = lib().orc_query_scores(self._p, seq, len(seq), out.ctypes.data)
        return k, out

    def query(self, seq: bytes, threshold: float, floor_mode: bool = False):
        """[(doc, score)] sorted (score desc, doc asc) -- the cobs result list."""
        k, sc = self.scores(seq, sliced=True)
        if k <= 0:
            return k, []
        t = threshold_terms(threshold, k, floor_mode)
        docs = np.nonzero(sc >= t)[0]
        order = np.lexsort((docs, -sc[docs].astype(np.int64)))
        return k, [(int(docs[i]), int(sc[docs[i]])) for i in order]

    def query_batch(self, seqs, threshold, threads=1, mode=0, floor_mode=False):
        """Timing driver; returns (total passing pairs, n_pass per query)."""
        cat = b"".join(seqs)
        offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(s) for s in seqs])
        n_pass = np.zeros(len(seqs), dtype=np.uint32)
        tot = lib().orc_query_batch(self._p, cat, offs.ctypes.data, len(seqs), threshold,
                                    int(floor_mode), threads, mode, n_pass.ctypes.data)
        if tot < 0:
            raise ValueError("oracle: invalid letter in a query")
        return tot, n_pass

    def query_text(self, records, threshold, floor_mode=False) -> str:
        """`cobs query` stdout for [(header_without_>, seq bytes)] ([A.8])."""
        names = self.doc_names
        out = []
        for header, seq in records:
            if len(seq) == 0:
                continue
            k, hits = self.query(seq, threshold, floor_mode)
            if k < 0:
                raise ValueError("oracle: invalid letter in query " + header)
            out.append(f"*{header}\t{len(hits)}\n")
            out.extend(f"{names[d]}\t{s}\n" for d, s in hits)
        return "".join(out)


def synth_genome(spec: SynthSpec, d: int) -> bytes:
    buf = C.create_string_buffer(spec.genome_len)
    lib().orc_synth_genome(C.byref(spec), d, buf)
    return buf.raw


def synth_read(specs, reads_seed: int, r: int, read_len: int, random_q8: int = 51,
               err_q16: int = 655) -> bytes:
    arr = (SynthSpec * len(specs))(*specs)
    buf = C.create_string_buffer(read_len)
    lib().orc_synth_read(arr, len(specs), reads_seed, r, read_len, random_q8, err_q16, buf)
    return buf.raw
