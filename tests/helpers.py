"""Shared test helpers (pure Python; no product imports)."""
import gzip
import lzma
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_BATCHES = ["aaa__01", "bbb__01", "bbb__02"]


def read_fasta(path):
    """[(header_without_>, seq)] with cobs's record rules ([A.8])."""
    recs = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            if line[0] in ">;":
                recs.append([line[1:], ""])
            elif recs:
                recs[-1][1] += line
    return [(h, s) for h, s in recs]


def golden_index_bytes(batch):
    return lzma.open(os.path.join(GOLDEN, f"{batch}.cobs_classic.xz")).read()


def golden_cobs_text(batch):
    return gzip.open(os.path.join(GOLDEN, f"{batch}.cobs.txt.gz")).read().decode()


def golden_match_text(batch, keep):
    return gzip.open(os.path.join(GOLDEN, f"n{keep}", f"{batch}____queries.gz")).read().decode()


def golden_filter_fa(keep):
    return open(os.path.join(GOLDEN, f"n{keep}", "queries.fa")).read()


def revcomp(s: str) -> str:
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def py_xxh64(data: bytes, seed: int = 0) -> int:
    """Pure-Python XXH64 (all lengths) -- third restatement used to brute-force tiny cases."""
    M = (1 << 64) - 1
    P1, P2, P3, P4, P5 = (0x9E3779B185EBCA87, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9,
                          0x85EBCA77C2B2AE63, 0x27D4EB2F165667C5)
    rotl = lambda x, r: ((x << r) | (x >> (64 - r))) & M
    rnd = lambda a, x: (rotl((a + x * P2) & M, 31) * P1) & M
    n = len(data)
    p = 0
    if n >= 32:
        v = [(seed + P1 + P2) & M, (seed + P2) & M, seed & M, (seed - P1) & M]
        while p <= n - 32:
            for i in range(4):
                v[i] = rnd(v[i], int.from_bytes(data[p:p + 8], "little"))
                p += 8
        h = (rotl(v[0], 1) + rotl(v[1], 7) + rotl(v[2], 12) + rotl(v[3], 18)) & M
        for i in range(4):
            h = ((h ^ rnd(0, v[i])) * P1 + P4) & M
    else:
        h = (seed + P5) & M
    h = (h + n) & M
    while p + 8 <= n:
        h ^= rnd(0, int.from_bytes(data[p:p + 8], "little"))
        h = (rotl(h, 27) * P1 + P4) & M
        p += 8
    if p + 4 <= n:
        h ^= (int.from_bytes(data[p:p + 4], "little") * P1) & M
        h = (rotl(h, 23) * P2 + P3) & M
        p += 4
    while p < n:
        h ^= (data[p] * P5) & M
        h = (rotl(h, 11) * P1) & M
        p += 1
    h ^= h >> 33
    h = (h * P2) & M
    h ^= h >> 29
    h = (h * P3) & M
    h ^= h >> 32
    return h


def brute_scores(docs, query: str, k=31, num_hashes=1, sig=None, canonicalize=True):
    """Scores by first principles: build bit sets in Python, query them."""
    canon = (lambda s: min(s, revcomp(s))) if canonicalize else (lambda s: s)
    rows = {}          # ONE signature array shared by all hash functions
    for d, doc in enumerate(docs):
        for i in range(len(doc) - k + 1):
            km = canon(doc[i:i + k]).encode()
            for j in range(num_hashes):
                rows.setdefault(py_xxh64(km, j) % sig, set()).add(d)
    scores = [0] * len(docs)
    for i in range(len(query) - k + 1):
        km = canon(query[i:i + k]).encode()
        hit = None
        for j in range(num_hashes):
            s = rows.get(py_xxh64(km, j) % sig, set())
            hit = s if hit is None else hit & s
        for d in hit:
            scores[d] += 1
    return scores
