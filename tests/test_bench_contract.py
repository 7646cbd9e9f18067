"""bench.py contract checks that need no GPU: workload arithmetic, placement, reference-arm JSON."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def ns(**kw):
    d = dict(workload="reads1k", db_scale=1.0, indexes=64, docs=4000, genome_len=1_000_000, reads=100_000,
             read_len=1000)
    d.update(kw)
    return argparse.Namespace(**d)


def test_config3_workload_numbers_match_baseline_md():
    w = bench.workload(ns())
    assert w["alg_bytes"] == 3_104_000_000_000          # 3.104 TB
    assert w["kmer_docs"] == 100_000 * 970 * 64 * 4000
    assert abs(w["signature_size"] / 999_970 - 2.8037) < 1e-3
    assert w["bases"] == 10 ** 8


def test_config4_workload_is_the_661k_database_shape():
    w = bench.workload(ns(workload="db661k"))
    assert w["n_indexes"] == 305 and w["row_bytes_per_kmer"] == 82741
    assert abs(w["alg_bytes"] / 1e12 - 8.03) < 0.01       # 8.03 TB per 1e8 bases
    total = sum(b["signature_size"] * ((b["n_docs"] + 7) // 8) for b in w["batches"])
    assert abs(total / 1.058455434059e12 - 1) < 1e-3      # decompressed database size
    placement, imbalance = bench.place(w, 8, 170 * 10 ** 9)
    assert sum(len(p) for p in placement) == 305 and imbalance < 1.02
    small = bench.workload(ns(workload="db661k", db_scale=0.05))
    assert len(bench.place(small, 1, 170 * 10 ** 9)[0][0]) == 305


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--reads", "120",
                        "--read-len", "150", "--indexes", "2", "--docs", "64", "--genome-len", "2000",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "bases/s" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    # both arms word the workload identically (run-dependent facts live outside `config`)
    w = bench.workload(ns(reads=120, read_len=150, indexes=2, docs=64, genome_len=2000))
    assert d["config"] == bench.config_dict(w)
    # files in -> files out: cobs_oracle | postprocess_cobs.py | gzip --fast, then filter_queries.py
    f = d["e2e_files"]
    assert f["value"] > 0 and f["unit"] == "bases/s"
    assert set(f["breakdown_s"]) >= {"match_pipeline_one_batch", "filter_queries_all_batches", "stage_cobs_query_alone"}
    assert "postprocess_cobs.py" in f["sample"] and f["outputs"]["filter_fasta_bytes"] > 0
    assert "set-up subprocess" in d["native_so_policy"]


def test_digest_helpers_are_placement_independent():
    """result_digest ingredients: per-batch digests keyed by the global batch rank, merged lists joined in
    query order from per-rank slices."""
    import numpy as np
    from phylign_b200.matcher import CAND_DT

    class FakeDist:
        world, rank = 2, 0
        def gather(self, obj):
            lo, hi, o, c = obj
            other = (hi, hi + 2, np.array([0, 1, 3], np.int64), np.array([(5, 1, 2, 3), (4, 0, 1, 1), (4, 1, 0, 0)], CAND_DT))
            return [other, obj]          # arrival order must not matter

    class FakeM:
        def merged_range(self):
            return 0, 2

    offs = np.array([0, 2, 3, 3, 3], np.uint64)
    cands = np.array([(9, 0, 0, 0), (8, 0, 1, 1), (7, 1, 0, 0)], CAND_DT)
    fo, fc = bench.full_merged(FakeDist(), FakeM(), offs, cands)
    assert fo.tolist() == [0, 2, 3, 4, 6] and fc["score"].tolist() == [9, 8, 7, 5, 4, 4]


def test_reference_arm_prefers_a_cobs_binary_on_path(tmp_path):
    """when an executable `cobs` is on PATH the CPU arm times IT (kind "reference"), for the
    query leg and inside the file pipeline.  Exercised with tests/fake_cobs.py (the oracle CLI posing as cobs)."""
    import oracle
    oracle.build()
    os.symlink(os.path.join(ROOT, "tests", "fake_cobs.py"), tmp_path / "cobs")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", ORC_CLI=oracle.CLI_PATH, PATH=f"{tmp_path}{os.pathsep}{os.environ['PATH']}")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--reads", "60",
                        "--read-len", "150", "--indexes", "2", "--docs", "64", "--genome-len", "2000",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cobs_on_path"] == str(tmp_path / "cobs")
    assert d["e2e_files"]["cobs_binary"] == str(tmp_path / "cobs") and d["value"] > 0
