"""Writers of intermediate/03_match/{batch}____{qfile}.gz (host side, on top of the library's
native writer phy_write_match_blocks).

What the reference produces per batch with
    cobs query ... | postprocess_cobs.py -n N | gzip --fast > {output.match}
(/root/reference/Snakefile:425-427,467-469,482-484) is written here for all resident indexes at
once, one query block at a time: formatting and zlib level 1 run on a pool of native threads (the
GIL is released inside the ctypes call), so a block can be written on a background Python thread
while the GPU works on the next one.  Files appear under their final names only on commit().
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .cobs_text import _cat


def default_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


class MatchFileSet:
    """One output file per index id, fed block by block in query order."""

    def __init__(self, paths: dict, indexes: dict, gzip_level: int = 1, threads: int = 0,
                 strip_prefix: bool = True):
        self._L = _lib.load()
        self.threads = threads or default_threads()
        self.strip_prefix = strip_prefix
        self.stats = _lib.WriteStats()
        self._files, self._keep = {}, []
        self.file_bytes = {}
        try:
            for idx_id, path in paths.items():
                h = C.c_void_p()
                _lib.check(self._L.phy_mfile_open(os.fsencode(path), gzip_level, C.byref(h)))
                self._files[idx_id] = h
        except Exception:
            self.abort()
            raise
        jobs = (_lib.MFileJob * max(1, len(self._files)))()
        for j, (idx_id, h) in enumerate(self._files.items()):
            ix = indexes[idx_id]
            if not hasattr(ix, "_names_cat"):
                ix._names_cat = _cat(ix.doc_names)
            ncat, noffs = ix._names_cat
            self._keep.append((ncat, noffs))
            jobs[j].file, jobs[j].idx_id, jobs[j].n_docs = h, idx_id, len(ix.doc_names)
            jobs[j].names, jobs[j].noffs = ncat, noffs.ctypes.data
        self._jobs = jobs

    def write_block(self, headers_cat, hoffs: np.ndarray, results_ptr, skip=None):
        """Append the cobs text of one query block to every file.  headers_cat: bytes or uint8 array with
        the header lines; hoffs[nq+1]: where the headers of this block's records start in it."""
        if not self._files:
            return
        if isinstance(headers_cat, np.ndarray):
            haddr = headers_cat.ctypes.data
        else:
            haddr = C.cast(C.c_char_p(headers_cat), C.c_void_p).value
        hoffs = np.ascontiguousarray(hoffs, dtype=np.uint64)
        _lib.check(self._L.phy_write_match_blocks(results_ptr, self._jobs, len(self._files), haddr,
                                                  hoffs.ctypes.data, None if skip is None else skip.ctypes.data,
                                                  int(self.strip_prefix), self.threads, C.byref(self.stats)))

    def commit(self):
        files, self._files = self._files, {}
        err = None
        for idx_id, h in files.items():
            n = C.c_uint64()
            rc = self._L.phy_mfile_commit(h, C.byref(n))
            if rc != 0 and err is None:
                err = _lib.PhylignCudaError(rc, self._L.phy_last_error(None).decode())
            self.file_bytes[idx_id] = n.value
        if err:
            raise err

    def abort(self):
        files, self._files = self._files, {}
        for h in files.values():
            self._L.phy_mfile_abort(h)

    def stats_dict(self) -> dict:
        s = self.stats
        return {"format_thread_s": s.format_s, "deflate_thread_s": s.deflate_s, "write_thread_s": s.write_s,
                "writer_wall_s": s.wall_s, "text_bytes": int(s.text_bytes), "file_bytes": int(s.file_bytes),
                "header_lines": int(s.n_header_lines), "hit_lines": int(s.n_hit_lines), "threads": self.threads}

    def __del__(self):
        try:
            self.abort()
        except Exception:
            pass
