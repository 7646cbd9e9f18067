"""Placement of batches (COBS indexes) on GPUs -- host logic, no GPU needed.

The reference parallelises over batches with one Snakemake job per batch, admitted by RAM
(`max_ram_mb`, /root/reference/Snakefile:401-407).  Here batches are placed on GPUs once:
longest-processing-time-first on the gather work per query k-mer (= row bytes, SURVEY.md
8(e)), subject to each GPU's HBM budget; what does not fit is returned as overflow rounds
that are streamed through after the resident set (config 4 with a capped budget).
"""
from __future__ import annotations

from dataclasses import dataclass, field


def row_stride(n_docs: int) -> int:
    """HBM bytes per row (must match stride_for() in csrc/capi.cu)."""
    rs = (n_docs + 7) // 8
    if rs <= 128:
        s = 16
        while s < rs:
            s <<= 1
        return s
    return (rs + 31) // 32 * 32


@dataclass
class Batch:
    name: str
    n_docs: int
    signature_size: int

    @property
    def work(self) -> int:            # bytes gathered per query k-mer (algorithmic)
        return (self.n_docs + 7) // 8

    @property
    def hbm_bytes(self) -> int:
        return self.signature_size * row_stride(self.n_docs) + 512 + 4 * self.n_docs


@dataclass
class Plan:
    n_ranks: int
    rounds: list = field(default_factory=list)   # rounds[r][rank] = [Batch, ...]

    def batches_of(self, rank: int, rnd: int = 0):
        return self.rounds[rnd][rank] if rnd < len(self.rounds) else []

    def work_of(self, rank: int) -> int:
        return sum(b.work for rnd in self.rounds for b in rnd[rank])

    @property
    def imbalance(self) -> float:
        w = [self.work_of(r) for r in range(self.n_ranks)]
        return max(w) / (sum(w) / len(w)) if sum(w) else 1.0


def assign(batches, n_ranks: int, hbm_budget: int) -> Plan:
    """LPT bin packing: heaviest batch first onto the least-loaded GPU that still has room."""
    plan = Plan(n_ranks)
    todo = sorted(batches, key=lambda b: (-b.work, -b.hbm_bytes, b.name))
    too_big = [b.name for b in todo if b.hbm_bytes > hbm_budget]
    if too_big:
        raise ValueError(f"batches larger than the per-GPU HBM budget: {too_big[:3]}")
    total_work = [0] * n_ranks          # across rounds: every GPU should gather the same bytes
    while todo:
        free = [hbm_budget] * n_ranks
        rnd = [[] for _ in range(n_ranks)]
        rest = []
        for b in todo:
            cands = [r for r in range(n_ranks) if free[r] >= b.hbm_bytes]
            if not cands:
                rest.append(b)
                continue
            r = min(cands, key=lambda r: (total_work[r], r))
            rnd[r].append(b)
            free[r] -= b.hbm_bytes
            total_work[r] += b.work
        plan.rounds.append(rnd)
        todo = rest
    return plan


def global_batch_ranks(names):
    """rank of each batch name in Python str order == the `batch` field of the merge key
    (/root/reference/scripts/filter_queries.py:135); identical on every rank."""
    return {b: i for i, b in enumerate(sorted(set(names)))}
