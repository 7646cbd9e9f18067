"""ctypes binding of libphylign_cuda.so (C ABI in include/phylign_cuda.h).

There is no CPU fallback: importing works anywhere (so host-side logic stays testable),
but every compute call raises PhylignCudaError when the library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libphylign_cuda.so")

PHY_OK = 0
ERR_NAMES = {-1: "PHY_ERR_CUDA", -2: "PHY_ERR_ARG", -3: "PHY_ERR_NOMEM", -4: "PHY_ERR_STATE",
             -5: "PHY_ERR_NCCL", -6: "PHY_ERR_QUERY", -7: "PHY_ERR_IO"}
NCCL_ID_BYTES = 128


class PhylignCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class MatchParams(C.Structure):
    _fields_ = [("threshold", C.c_double), ("top_n", C.c_uint32), ("floor_mode", C.c_uint32)]


class Hit(C.Structure):
    _fields_ = [("doc", C.c_uint32), ("score", C.c_uint32)]


class Unit(C.Structure):
    _fields_ = [("query", C.c_uint32), ("index", C.c_uint32), ("n_pass", C.c_uint32),
                ("n_kept", C.c_uint32), ("offset", C.c_uint64)]


class Results(C.Structure):
    _fields_ = [("n_queries", C.c_uint32), ("n_indexes", C.c_uint32), ("n_units", C.c_uint64),
                ("units", C.POINTER(Unit)), ("n_hits", C.c_uint64), ("hits", C.POINTER(Hit)),
                ("n_kmers", C.POINTER(C.c_uint32)), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64)]


class Cand(C.Structure):
    _fields_ = [("score", C.c_uint32), ("batch_rank", C.c_uint32), ("doc", C.c_uint32),
                ("ref_rank", C.c_uint32)]


class Merged(C.Structure):
    _fields_ = [("n_queries", C.c_uint32), ("offs", C.POINTER(C.c_uint64)),
                ("cands", C.POINTER(Cand)), ("d2h_bytes", C.c_uint64)]


class MatchText(C.Structure):
    _fields_ = [("n_blocks", C.c_uint64), ("n_hits", C.c_uint64), ("q_off", C.POINTER(C.c_uint64)),
                ("q_len", C.POINTER(C.c_uint32)), ("first_hit", C.POINTER(C.c_uint64)),
                ("ref_off", C.POINTER(C.c_uint64)), ("ref_len", C.POINTER(C.c_uint32)),
                ("kmers", C.POINTER(C.c_uint32))]


class Fasta(C.Structure):
    _fields_ = [("n", C.c_uint32), ("simple", C.c_uint8), ("seqs_pinned", C.c_uint8),
                ("seqs", C.c_void_p), ("soffs", C.POINTER(C.c_uint64)),
                ("headers", C.c_void_p), ("hoffs", C.POINTER(C.c_uint64)),
                ("name_len", C.POINTER(C.c_uint32))]


class MFileJob(C.Structure):
    _fields_ = [("file", C.c_void_p), ("idx_id", C.c_uint32), ("n_docs", C.c_uint32),
                ("names", C.c_char_p), ("noffs", C.c_void_p)]


class WriteStats(C.Structure):
    _fields_ = [("format_s", C.c_double), ("deflate_s", C.c_double), ("write_s", C.c_double),
                ("wall_s", C.c_double), ("text_bytes", C.c_uint64), ("file_bytes", C.c_uint64),
                ("n_header_lines", C.c_uint64), ("n_hit_lines", C.c_uint64)]


class IndexInfo(C.Structure):
    _fields_ = [("signature_size", C.c_uint64), ("num_hashes", C.c_uint64), ("hbm_bytes", C.c_uint64),
                ("term_size", C.c_uint32), ("n_docs", C.c_uint32), ("row_size", C.c_uint32),
                ("row_stride", C.c_uint32), ("batch_rank", C.c_uint32), ("canonicalize", C.c_uint8),
                ("committed", C.c_uint8)]


class SynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_docs", C.c_uint32), ("genome_len", C.c_uint32),
                ("clade_size", C.c_uint32), ("clade_sub_q16", C.c_uint32), ("doc_sub_q16", C.c_uint32)]


# every symbol include/phylign_cuda.h declares: name -> (restype, argtypes)
_P = C.c_void_p
PROTOTYPES = {
    "phy_abi_version": (C.c_int, []),
    "phy_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "phy_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_uint64]),
    "phy_ctx_destroy": (None, [_P]),
    "phy_last_error": (C.c_char_p, [_P]),
    "phy_index_begin": (C.c_int, [_P, C.c_char_p, C.c_uint32, C.c_uint8, C.c_uint64, C.c_uint64,
                                  C.c_uint32, C.POINTER(C.c_int)]),
    "phy_index_push": (C.c_int, [_P, C.c_int, C.c_void_p, C.c_uint64]),
    "phy_index_load_file": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_uint64, C.c_int]),
    "phy_index_commit": (C.c_int, [_P, C.c_int]),
    "phy_index_evict": (C.c_int, [_P, C.c_int]),
    "phy_index_set_ranks": (C.c_int, [_P, C.c_int, C.c_uint32, C.c_void_p]),
    "phy_index_set_active": (C.c_int, [_P, C.c_int, C.c_int]),
    "phy_index_info_get": (C.c_int, [_P, C.c_int, C.POINTER(IndexInfo)]),
    "phy_index_count": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "phy_index_download": (C.c_int, [_P, C.c_int, C.c_void_p, C.c_uint64]),
    "phy_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "phy_host_free": (None, [C.c_void_p]),
    "phy_queries_set": (C.c_int, [_P, C.c_void_p, C.c_void_p, C.c_uint32]),
    "phy_fix_bases": (C.c_int, [_P, C.c_void_p, C.c_uint64]),
    "phy_match_run": (C.c_int, [_P, C.POINTER(MatchParams), C.c_uint32]),
    "phy_results_fetch": (C.c_int, [_P, C.POINTER(C.POINTER(Results))]),
    "phy_results_free": (None, [C.POINTER(Results)]),
    "phy_match": (C.c_int, [_P, C.POINTER(MatchParams), C.POINTER(C.POINTER(Results))]),
    "phy_scores": (C.c_int, [_P, C.c_int, C.c_void_p]),
    "phy_merged_fetch": (C.c_int, [_P, C.POINTER(C.POINTER(Merged))]),
    "phy_merged_free": (None, [C.POINTER(Merged)]),
    "phy_merged_range": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "phy_merge_host": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "phy_format_cobs_text": (C.c_int, [C.POINTER(Results), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p,
                                      C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "phy_format_filter_fasta": (C.c_int, [C.POINTER(Merged), C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_uint64)]),
    "phy_text_free": (None, [C.c_void_p]),
    "phy_parse_match_text": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.POINTER(MatchText))]),
    "phy_match_text_free": (None, [C.POINTER(MatchText)]),
    "phy_fasta_read": (C.c_int, [C.c_char_p, C.POINTER(C.POINTER(Fasta))]),
    "phy_fasta_free": (None, [C.POINTER(Fasta)]),
    "phy_write_filter_fasta": (C.c_int, [C.c_char_p, C.POINTER(Merged), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]),
    "phy_mfile_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "phy_mfile_commit": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "phy_mfile_abort": (None, [C.c_void_p]),
    "phy_write_match_blocks": (C.c_int, [C.POINTER(Results), C.POINTER(MFileJob), C.c_uint32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_int, C.POINTER(WriteStats)]),
    "phy_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "phy_nccl_init": (C.c_int, [_P, C.c_void_p, C.c_int, C.c_int]),
    "phy_nccl_finalize": (C.c_int, [_P]),
    "phy_timer_start": (C.c_int, [_P]),
    "phy_timer_stop": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "phy_sync": (C.c_int, [_P]),
    "phy_last_phase_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "phy_last_gather_bytes": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "phy_last_gather_bytes_of": (C.c_int, [_P, C.c_int, C.POINTER(C.c_uint64)]),
    "phy_ctx_budget": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "phy_ctx_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "phy_flush_l2": (C.c_int, [_P]),
    "phy_index_synth": (C.c_int, [_P, C.c_int, C.POINTER(SynthSpec)]),
    "phy_index_insert": (C.c_int, [_P, C.c_int, C.c_void_p]),
    "phy_synth_reads": (C.c_int, [_P, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32,
                                  C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
}

_lib = None


def load():
    """dlopen the CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PhylignCudaError(-1, f"{LIB_PATH} is missing: build it with "
                                       "`python -m phylign_b200.build` (nvcc, sm_100a)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.phy_abi_version() != 1:
            raise PhylignCudaError(-2, "ABI version mismatch")
        _lib = L
    return _lib


def check(code, ctx=None):
    if code != PHY_OK:
        msg = load().phy_last_error(ctx)
        raise PhylignCudaError(code, msg.decode() if msg else "")
