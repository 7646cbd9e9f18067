"""Host-side multi-GPU logic on CPU: placement over the real database shape, and the
world_size-2 control plane over gloo."""
import os
import socket
import subprocess
import sys

import pytest

from phylign_b200 import sharding
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def db_batches():
    out = []
    for line in open(os.path.join(H.GOLDEN, "db_shape.tsv")):
        if line.startswith("#"):
            continue
        name, nbytes, docs = line.split("\t")
        docs = int(docs)
        body = int(nbytes)          # header is <0.01% of the file; good enough for placement
        out.append(sharding.Batch(name, docs, body // ((docs + 7) // 8)))
    return out


def test_db_shape_fixture_matches_survey_appendix_d():
    b = db_batches()
    assert len(b) == 305 and sum(x.n_docs for x in b) == 661405
    assert sum(x.work for x in b) == 82741                       # bytes gathered per query k-mer
    assert sum(1 for x in b if x.n_docs == 4000) == 141


def test_whole_db_fits_8_gpus_and_is_balanced():
    b = db_batches()
    plan = sharding.assign(b, 8, 170 * 10 ** 9)
    assert len(plan.rounds) == 1                                  # 1.06 TB (+2.4% padding) resident in 8 x HBM
    placed = [x.name for r in range(8) for x in plan.batches_of(r)]
    assert sorted(placed) == sorted(x.name for x in b)
    assert plan.imbalance < 1.02
    for r in range(8):
        assert sum(x.hbm_bytes for x in plan.batches_of(r)) <= 170 * 10 ** 9


def test_capped_budget_streams_overflow_rounds():
    b = db_batches()
    plan = sharding.assign(b, 8, 60 * 10 ** 9)                    # force streaming (config 4 overflow)
    assert len(plan.rounds) >= 2
    placed = [x.name for rnd in plan.rounds for rank in rnd for x in rank]
    assert sorted(placed) == sorted(x.name for x in b) and len(set(placed)) == 305
    for rnd in plan.rounds:
        for rank in rnd:
            assert sum(x.hbm_bytes for x in rank) <= 60 * 10 ** 9
    assert plan.imbalance < 1.05
    with pytest.raises(ValueError):
        sharding.assign(b, 8, 10 ** 9)


def test_stride_rule_matches_library():
    for docs, want in ((1, 16), (128, 16), (129, 32), (200, 32), (257, 64), (664, 128), (520, 128), (4000, 512),
                       (4097, 544)):
        assert sharding.row_stride(docs) == want


def test_global_batch_ranks_are_string_order():
    r = sharding.global_batch_ranks(["b__02", "a__10", "b__01", "B__01"])
    assert r == {"B__01": 0, "a__10": 1, "b__01": 2, "b__02": 3}


WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from phylign_b200 import sharding
from tests.test_sharding import db_batches
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
# rank 0 makes the id (stand-in for phy_nccl_unique_id), everyone must end with the same bytes
obj = [bytes(range(128)) if rank == 0 else None]
dist.broadcast_object_list(obj, src=0)
plan = sharding.assign(db_batches(), world, 600 * 10 ** 9)
mine = [b.name for b in plan.batches_of(rank)]
ranks = sharding.global_batch_ranks([b.name for b in db_batches()])
allnames = [None] * world
dist.all_gather_object(allnames, mine)
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
dist.barrier()
print(json.dumps({"rank": rank, "id_ok": obj[0] == bytes(range(128)), "n": len(mine),
                  "all": sorted(sum(allnames, [])) == sorted(ranks), "max": float(t[0]),
                  "rank_of_first": ranks[mine[0]]}))
'''


def test_world_size_2_control_plane_over_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), PYTHONPATH=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=120)
        assert p.returncode == 0, e[-2000:]
        outs.append(__import__("json").loads(o.strip().splitlines()[-1]))
    assert all(o["id_ok"] and o["all"] and o["max"] == 2.0 for o in outs)
    assert outs[0]["n"] + outs[1]["n"] == 305
