#!/usr/bin/env python3
"""bench.py -- COBS-match throughput of the B200 path (and of the CPU oracle beside it).

Workload (BASELINE.json configs[2]): 100k synthetic 1 kbp reads vs 64 synthetic batch
indexes (4000 docs x 1 Mbp genomes each, k=31, 1 hash, fpr 0.3 -> ~1.43 GB per index)
resident in HBM, threshold 0.7, top-N 100 + ties, per-query merge over all indexes.
A step = one pass of all reads over all indexes.  N GPUs: the 64 indexes are sharded
round-robin over the ranks (strong scaling, fixed total work), queries replicated, per-GPU
candidate lists gathered with NCCL and merged on rank 0 inside the step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "COBS-match query bases/s over whole index"
UNIT = "bases/s"
THRESHOLD, TOP_N = 0.7, 100
READS_SEED, RANDOM_Q8, ERR_Q16 = 3, 51, 655


def workload(args):
    """Batches (indexes) + reads of the selected workload.

    reads1k (default, BASELINE configs[2]): --indexes equal synthetic batches.
    db661k (BASELINE configs[3]): the 305 batches of the 661k database, document counts and
    decompressed sizes from the reference's data files (phylign_b200/data/db_shape.tsv),
    content synthetic; --db-scale shrinks every signature_size (1.0 = the real 1.06 TB)."""
    ratio = -1.0 / math.log(1.0 - 0.3)                      # cobs classic-construct: fpr 0.3, 1 hash
    batches = []
    if args.workload == "db661k":
        for line in open(os.path.join(ROOT, "phylign_b200", "data", "db_shape.tsv")):
            if line.startswith("#"):
                continue
            name, nbytes, docs = line.split("\t")
            docs = int(docs)
            sig = max(4096, int(int(nbytes) / ((docs + 7) // 8) * args.db_scale))
            batches.append(dict(name=name, n_docs=docs, signature_size=sig,
                                genome_len=max(1200, int(sig / ratio)), seed=1000 + len(batches)))
    else:
        sig = int(math.ceil((args.genome_len - 30) * ratio))
        for i in range(args.indexes):
            batches.append(dict(name=batch_name(i), n_docs=args.docs, signature_size=sig,
                                genome_len=args.genome_len, seed=1000 + i))
    w = dict(batches=batches, n_indexes=len(batches), n_docs=max(b["n_docs"] for b in batches),
             n_reads=args.reads, read_len=args.read_len, name=args.workload)
    w["signature_size"] = batches[0]["signature_size"]
    w["row_bytes_per_kmer"] = sum((b["n_docs"] + 7) // 8 for b in batches)
    w["kmers_per_read"] = max(w["read_len"] - 30, 0)
    w["bases"] = w["n_reads"] * w["read_len"]
    # algorithmic bytes = sum_q K_q * sum_b h_b * ceil(D_b/8)
    w["alg_bytes"] = w["n_reads"] * w["kmers_per_read"] * w["row_bytes_per_kmer"]
    w["kmer_docs"] = w["n_reads"] * w["kmers_per_read"] * sum(b["n_docs"] for b in batches)
    return w


def spec_kwargs(b, w=None):
    if isinstance(b, int):                                   # index number of the reads1k workload
        b = w["batches"][b]
    return dict(seed=b["seed"], n_docs=b["n_docs"], genome_len=b["genome_len"], clade_size=32,
                clade_sub_q16=328, doc_sub_q16=328)


def batch_name(i):
    return f"synth_species_{i:03d}__01"


def place(w, world, budget):
    """LPT placement of the batches on the ranks (phylign_b200/sharding.py); one resident round."""
    from phylign_b200 import sharding
    bl = [sharding.Batch(b["name"], b["n_docs"], b["signature_size"]) for b in w["batches"]]
    plan = sharding.assign(bl, world, budget)
    if len(plan.rounds) != 1:
        raise SystemExit(f"workload needs {len(plan.rounds)} resident rounds on {world} GPU(s): "
                         "use more GPUs or --db-scale")
    by_name = {b["name"]: b for b in w["batches"]}
    return [[by_name[x.name] for x in plan.batches_of(r)] for r in range(world)], plan.imbalance


def config_dict(w, extra=None):
    if w["name"] == "db661k":
        what = (f"{w['n_indexes']} synthetic COBS classic indexes shaped like the 661k database "
                f"(docs and sizes per batch from data/decompressed_indexes_sizes.txt + 661k_batches.txt, "
                f"{w['row_bytes_per_kmer']} B gathered per k-mer) (BASELINE.json configs[3])")
    else:
        what = (f"{w['n_indexes']} synthetic COBS classic indexes ({w['n_docs']} docs x "
                f"{w['batches'][0]['genome_len']} bp, k=31, h=1, fpr=0.3) (BASELINE.json configs[2])")
    c = {"workload": f"{w['n_reads']} synthetic {w['read_len']} bp reads vs {what} resident in HBM, "
                     f"-t {THRESHOLD}, top-{TOP_N}+ties, cross-index merge",
         "n_reads": w["n_reads"], "read_len": w["read_len"], "n_indexes": w["n_indexes"],
         "docs_per_index": w["n_docs"], "threshold": THRESHOLD, "top_n": TOP_N,
         "algorithmic_bytes_per_step": w["alg_bytes"], "kmer_docs_per_step": w["kmer_docs"],
         "l2_hygiene": "inputs (index shard) exceed the 126 MB L2; no flush needed"}
    if extra:
        c.update(extra)
    return c


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------- CPU arm
def oracle_index_from_device(m, idx_id, w):
    """Wrap a device-built synthetic index as an oracle index in host RAM (setup, untimed)."""
    import oracle
    oidx = oracle.OracleIndex.new(w["batches"][0]["n_docs"], w["batches"][0]["signature_size"])
    body = m.download_index(idx_id)
    oidx.body.reshape(-1)[:] = np.frombuffer(body, dtype=np.uint8)
    return oidx


def time_oracle(oidx, reads, threads):
    """Seconds for one pass of `reads` over ONE index, best of the two thread layouts."""
    out = {}
    for mode, name in ((0, "cobs-shape: queries serial, 128-doc slices over threads"),
                       (1, "queries over threads")):
        t0 = time.perf_counter()
        oidx.query_batch(reads, THRESHOLD, threads=threads, mode=mode)
        out[name] = time.perf_counter() - t0
    best = min(out, key=out.get)
    return out[best], best, out


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    import oracle
    oracle.build()
    cores = host_cores()
    n_sample = min(w["n_reads"], args.cpu_sample_reads)
    specs_kw = [spec_kwargs(b) for b in w["batches"]]
    w0 = w["batches"][0]
    try:
        from phylign_b200 import _lib
        from phylign_b200.matcher import Matcher
        m = Matcher(int(os.environ.get("LOCAL_RANK", 0)))
        i0 = m.add_synth_index(w0["name"], _lib.SynthSpec(**specs_kw[0]), w0["signature_size"])
        oidx = oracle_index_from_device(m, i0, w)
        raw = m.synth_reads([_lib.SynthSpec(**k) for k in specs_kw], READS_SEED, 0, n_sample, w["read_len"],
                            RANDOM_Q8, ERR_Q16)
        m.close()
        built = "index 0 and the reads synthesised on the GPU (setup only), then copied to host RAM"
    except Exception as e:  # no GPU: build the (small) workload with the oracle itself
        if w0["n_docs"] * w0["genome_len"] > 5e7:
            print(json.dumps({"impl": "reference", "unavailable": f"cannot synthesise the index without a GPU: {e}"}))
            return
        ospecs = [oracle.SynthSpec(**k) for k in specs_kw]
        oidx = oracle.OracleIndex.construct([oracle.synth_genome(ospecs[0], d) for d in range(w0["n_docs"])],
                                            signature_size_override=w0["signature_size"])
        raw = b"".join(oracle.synth_read(ospecs, READS_SEED, r, w["read_len"], RANDOM_Q8, ERR_Q16)
                       for r in range(n_sample))
        built = "index 0 and the reads built by the oracle on the CPU"
    L = w["read_len"]
    reads = [raw[r * L:(r + 1) * L] for r in range(n_sample)]
    _, mode_name, _ = time_oracle(oidx, reads[:max(1, n_sample // 8)], cores)  # pick the faster layout
    mode = 0 if mode_name.startswith("cobs") else 1
    for _ in range(args.warmup):
        oidx.query_batch(reads[:max(1, n_sample // 8)], THRESHOLD, threads=cores, mode=mode)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oidx.query_batch(reads, THRESHOLD, threads=cores, mode=mode)
    dt = (time.perf_counter() - t0) / args.steps
    # one index timed; the whole database is n_indexes such passes (independent batches)
    value = n_sample * L / (dt * w["n_indexes"])
    sample = (f"{n_sample} of the {w['n_reads']} reads against 1 of the {w['n_indexes']} indexes per step "
              f"({dt:.2f} s), extrapolated x{w['n_indexes']} indexes; CPU restatement of cobs 0.2.1 classic "
              f"query (oracle port, NOT the cobs binary), layout '{mode_name}'; {built}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": config_dict(w),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- GPU arm
def run_ours(args, w, rank, world, local_rank):
    from phylign_b200 import _lib
    from phylign_b200.matcher import Matcher, nccl_unique_id
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("gloo", rank=rank, world_size=world)  # control plane only

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    m = Matcher(local_rank)
    if world > 1:
        obj = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        m.nccl_init(obj[0], rank, world)
    specs = [_lib.SynthSpec(**spec_kwargs(b)) for b in w["batches"]]
    placement, imbalance = place(w, world, args.hbm_budget_gb * 10 ** 9)
    t_build = time.perf_counter()
    local = placement[rank]
    ids = {b["name"]: m.add_synth_index(b["name"], _lib.SynthSpec(**spec_kwargs(b)), b["signature_size"])
           for b in local}
    m.sync()
    t_build = time.perf_counter() - t_build
    m.set_ranks([b["name"] for b in w["batches"]])
    L = w["read_len"]
    raw = m.synth_reads(specs, READS_SEED, 0, w["n_reads"], L, RANDOM_Q8, ERR_Q16)
    offs = np.arange(w["n_reads"] + 1, dtype=np.uint64) * L
    # the step's inputs live in pinned host memory (phy_host_alloc), as the e2e contract asks
    from phylign_b200.matcher import PinnedBuffer
    pin_raw, pin_offs = PinnedBuffer(len(raw)), PinnedBuffer(offs.nbytes)
    pin_raw.array[:] = np.frombuffer(raw, dtype=np.uint8)
    pin_offs.array[:] = offs.view(np.uint8)
    local_alg_bytes = w["n_reads"] * w["kmers_per_read"] * sum((b["n_docs"] + 7) // 8 for b in local)

    # ---- device-resident throughput (`value`): queries already in HBM
    m.set_queries_raw(raw, offs)
    for _ in range(args.warmup):
        m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    if dist is not None:
        import torch
        torch.cuda.synchronize()
    m.sync()
    barrier()
    launches, gather_ms, phases, gathered = 0, [], [], 0
    m.timer_start()
    for _ in range(args.steps):
        m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
        gathered = m.gathered_bytes()
        ph = m.phase_ms()
        gather_ms.append(ph[1])
        phases.append(ph[:3])
        launches += int(ph[3])
    dev_ms = m.timer_stop()
    m.sync()
    barrier()
    clocks = sampler.stop() if sampler else None
    dev_ms = max_over_ranks(dev_ms)
    ms_per_step = dev_ms / args.steps
    value = w["bases"] / (ms_per_step * 1e-3)

    # ---- end to end through the public API: host buffers in, host results out, every step
    for _ in range(max(2, args.warmup)):   # warm-up: the pinned result pool reaches its steady state after 2 passes
        m.set_queries_raw(pin_raw, pin_offs)
        m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
        res = m.fetch()               # held like in the timed loop, so the pool ends up with both
        moffs, mcands = m.merged()    # generations of result blocks before timing starts
    m.sync()
    barrier()
    h2d = d2h = 0
    parts = np.zeros(4)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ta = time.perf_counter()
        m.set_queries_raw(pin_raw, pin_offs)               # H2D of the step's inputs (pinned host buffers)
        tb = time.perf_counter()
        m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
        tc = time.perf_counter()
        res = m.fetch()                                    # D2H: per-(query,index) hit lists (03_match content)
        td = time.perf_counter()
        moffs, mcands = m.merged()                         # D2H: merged top-N lists (04_filter content)
        te = time.perf_counter()
        parts += [tb - ta, tc - tb, td - tc, te - td]
        h2d = len(raw) + offs.nbytes
        d2h = res.d2h_bytes + moffs.nbytes + mcands.nbytes
    m.sync()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    barrier()
    e2e_value = w["bases"] / e2e_s

    # ---- the same gather with the exact threshold pruning switched off (explains the roofline)
    m.set_option("prune", 0)
    unpruned_ms = []
    for i in range(2 + 3):
        m.match_run(THRESHOLD, TOP_N, merge_top_n=TOP_N)
        if i >= 2:
            unpruned_ms.append(m.phase_ms()[1])
    m.set_option("prune", 1)
    barrier()
    if rank != 0:
        m.close()
        return
    peak, peak_src = measured_peak()
    kernel = ("gather_count_ring_kernel<32,10,3,4>" if w["name"] == "reads1k" else
              "gather_count_ring_kernel<LPR,10,3,4> (one launch per row-width class)")
    traffic = None
    try:   # DRAM bytes per launch from the committed ncu capture of this exact workload, else null
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = f"{kernel}|{w['name']}|{w['n_reads']}|{w['read_len']}|{w['n_indexes']}|{w['n_docs']}|{world}|prune1"
        traffic = tj[key]["traffic_bytes"] if key in tj else None
    except Exception:
        pass
    g_ms = float(np.mean(gather_ms))
    achieved = local_alg_bytes / (g_ms * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": config_dict(w, {"sharding": f"indexes placed LPT-by-row-bytes on {world} GPU(s) "
                                                  f"(imbalance {imbalance:.3f}), queries replicated",
                                      "index_build_s": round(t_build, 2),
                                      "kmer_docs_per_s": w["kmer_docs"] / (ms_per_step * 1e-3),
                                      "phase_ms_hash_gather_merge": [round(float(x), 3) for x in np.mean(phases, axis=0)],
                                      "pruning": "exact threshold pruning on (output-identical); "
                                                 "roofline.unpruned = the same step with every row read",
                                      "n_units": int(len(res.units)), "n_hits": int(len(res.hits)),
                                      "n_merged": int(len(mcands))}),
            # units one launch processes = (k-mer, index) pairs whose row was gathered; the exact
            # threshold pruning ends a (query,index) unit once no document can reach -t any more, so
            # that is fewer pairs than K_q x indexes.  achieved = gathered row bytes / time.
            "roofline": {"bound": "hbm", "achieved": gathered / (g_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": gathered / (g_ms * 1e-3) / 1e9 / peak, "traffic": traffic,
                         "kernel": kernel,
                         "bytes_per_launch": int(gathered),
                         "all_pairs": {"bytes_per_launch": int(local_alg_bytes), "achieved": achieved,
                                       "frac": achieved / peak,
                                       "what": "SURVEY 8(d) bytes of ALL (k-mer,index) pairs of the step / the same "
                                               "time: > peak because pairs that cannot change the output are "
                                               "never gathered (results are bit-identical, tests/ run with "
                                               "pruning on)"},
                         "unpruned": {"ms": float(np.mean(unpruned_ms)),
                                      "achieved": local_alg_bytes / (float(np.mean(unpruned_ms)) * 1e-3) / 1e9,
                                      "frac": local_alg_bytes / (float(np.mean(unpruned_ms)) * 1e-3) / 1e9 / peak,
                                      "what": "same launch with pruning off: every pair gathered"},
                         "note": f"row bytes gathered by rank 0's launch / mean CUDA-event duration of the "
                                 f"gather phase ({g_ms:.2f} ms, one launch per step); peak = {peak_src}"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3,
                    "breakdown_ms_set_run_fetch_merged": [round(float(x) * 1e3 / args.steps, 2) for x in parts]},
            "gpu_launches": launches, "clocks": clocks}
    # ---- CPU baseline beside it (N=1 only): the oracle on the host cores, bounded sample
    if world == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            oracle.build()
            cores = host_cores()
            n_sample = min(w["n_reads"], args.cpu_sample_reads)
            oidx = oracle_index_from_device(m, ids[w["batches"][0]["name"]], w)
            reads = [raw[r * L:(r + 1) * L] for r in range(n_sample)]
            dt, mode_name, both = time_oracle(oidx, reads, cores)
            v = n_sample * L / (dt * w["n_indexes"])
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{n_sample} reads against 1 of the {w['n_indexes']} indexes ({dt:.2f} s), extrapolated "
                          f"x{w['n_indexes']}; oracle port of cobs 0.2.1 classic query (NOT the cobs binary), "
                          f"layout '{mode_name}'; both layouts: " +
                          ", ".join(f"{k}: {t:.2f} s" for k, t in both.items())}
        except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port",
                                    "sample": f"failed: {e}"}
    m.close()
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100_000)
    ap.add_argument("--read-len", type=int, default=1000)
    ap.add_argument("--indexes", type=int, default=64)
    ap.add_argument("--docs", type=int, default=4000)
    ap.add_argument("--genome-len", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-reads", type=int, default=20_000)
    ap.add_argument("--workload", default="reads1k", choices=["reads1k", "db661k"])
    ap.add_argument("--db-scale", type=float, default=1.0, help="db661k: scale every signature_size")
    ap.add_argument("--hbm-budget-gb", type=float, default=170.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    args.gpus = max(args.gpus, world)
    w = workload(args)
    if args.impl == "reference":
        run_reference(args, w, rank, world)
    else:
        run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
