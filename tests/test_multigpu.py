"""Two-GPU parity of the NCCL merge (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, time
root, rank, world, idfile, keep, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]), sys.argv[6]
sharded = len(sys.argv) > 7 and sys.argv[7] == "sharded"
sys.path.insert(0, root)
from phylign_b200.matcher import Matcher, nccl_unique_id
from phylign_b200.cobs_index import ref_of
from tests.text_twins import format_filter_fasta
from tests import helpers as H
if rank == 0:
    with open(idfile + ".tmp", "wb") as f: f.write(nccl_unique_id())
    os.replace(idfile + ".tmp", idfile)
while not os.path.exists(idfile): time.sleep(0.05)
m = Matcher(rank)
m.nccl_init(open(idfile, "rb").read(), rank, world)
mine = [b for i, b in enumerate(H.GOLDEN_BATCHES) if i % world == rank]
for b in mine: m.load_index(os.path.join(H.GOLDEN, b + ".cobs_classic.xz"))
m.set_ranks(H.GOLDEN_BATCHES)
qs = H.read_fasta(os.path.join(H.GOLDEN, "queries.fa"))
if sharded:     # every rank finalises its slice of the queries; bases uploaded 1/R per rank + all-gather
    m.set_option("merge_mode", 1)
    m.set_option("shard_query_upload", 1)
    qs = qs * 400          # > 1 MB of bases, so the sharded upload path is taken
m.set_queries(qs)
m.match_run(0.7, top_n=keep, merge_top_n=keep)
offs, cands = m.merged()
if sharded:
    import numpy as np, pickle
    lo, hi = m.merged_range()
    assert (lo, hi) == (len(qs) * rank // world, len(qs) * (rank + 1) // world)
    o = offs.astype(np.int64)
    assert o[lo] == 0 and o[hi] == len(cands) and (np.diff(o[:lo + 1]) == 0).all() and (np.diff(o[hi:]) == 0).all()
    pickle.dump((lo, hi, o[lo:hi + 1] - o[lo], np.array(cands)), open(out + f".part{rank}", "wb"))
    m.close()
    sys.exit(0)
if rank == 0:
    import lzma
    from phylign_b200.cobs_index import parse_bytes
    refs = {}
    for r, b in enumerate(sorted(H.GOLDEN_BATCHES)):
        hdr, _ = parse_bytes(H.golden_index_bytes(b))
        refs[r] = [ref_of(n) for n in hdr.doc_names]
    open(out, "w").write(format_filter_fasta([(h.split(" ")[0], s) for h, s in qs], offs, cands, refs))
else:
    assert len(cands) == 0
m.close()
'''


@pytest.mark.parametrize("keep", [1, 100])
def test_two_gpu_nccl_merge_equals_reference_filter(tmp_path, keep):
    import ctypes as C
    from phylign_b200 import _lib
    n = C.c_int()
    _lib.load().phy_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    out = tmp_path / "out.fa"
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", str(tmp_path / "id"), str(keep), str(out)],
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    for p in procs:
        _, e = p.communicate(timeout=300)
        assert p.returncode == 0, e[-3000:]
    assert out.read_text() == H.golden_filter_fa(keep)


@pytest.mark.parametrize("keep", [1, 100])
def test_two_gpu_query_sharded_merge_and_sharded_upload(tmp_path, keep):
    """merge_mode 1 + shard_query_upload: the union of the ranks' slices == the reference filter output
    (the golden query set repeated 400 times: every repetition must give the golden answer)."""
    import ctypes as C
    import pickle
    import numpy as np
    from phylign_b200 import _lib
    from phylign_b200.cobs_index import parse_bytes, ref_of
    from tests.text_twins import format_filter_fasta
    n = C.c_int()
    _lib.load().phy_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    out = tmp_path / "out.fa"
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", str(tmp_path / "id"), str(keep), str(out),
                               "sharded"], stderr=subprocess.PIPE, text=True) for r in range(2)]
    for p in procs:
        _, e = p.communicate(timeout=300)
        assert p.returncode == 0, e[-3000:]
    qs = H.read_fasta(os.path.join(H.GOLDEN, "queries.fa"))
    refs = {}
    for r, b in enumerate(sorted(H.GOLDEN_BATCHES)):
        hdr, _ = parse_bytes(H.golden_index_bytes(b))
        refs[r] = [ref_of(nm) for nm in hdr.doc_names]
    want = H.golden_filter_fa(keep)
    got = []
    for r in range(2):
        lo, hi, o, cands = pickle.load(open(str(out) + f".part{r}", "rb"))
        recs = [(qs[q % len(qs)][0].split(" ")[0], qs[q % len(qs)][1]) for q in range(lo, hi)]
        got.append(format_filter_fasta(recs, o.astype(np.uint64), cands, refs))
    assert "".join(got) == want * 400


def test_match_db_gpus_2_writes_all_outputs(tmp_path):
    """`match-db --gpus 2`: worker per GPU, NCCL merge, several rounds and query blocks (so one rank
    sits out some merges with no index): match files and 04_filter equal the golden files."""
    import ctypes as C
    import gzip
    from phylign_b200 import _lib
    n = C.c_int()
    _lib.load().phy_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    batches = tmp_path / "batches.txt"
    batches.write_text("\n".join(H.GOLDEN_BATCHES) + "\n")
    mdir, out = tmp_path / "03_match", tmp_path / "04" / "q.fa"
    r = subprocess.run([sys.executable, "-m", "phylign_b200.cli", "match-db", "--cobs-dir", H.GOLDEN, "--batches",
                        str(batches), "-q", os.path.join(H.GOLDEN, "queries.fa"), "--match-dir", str(mdir),
                        "--filter-out", str(out), "-t", "0.7", "-n", "3", "--gpus", "2", "--round-bytes", "500000",
                        "--query-block-bases", "4000"], capture_output=True, text=True, cwd=ROOT,
                       env=dict(os.environ, PYTHONPATH=ROOT), timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    for b in H.GOLDEN_BATCHES:
        assert gzip.open(mdir / f"{b}____queries.gz", "rt").read() == H.golden_match_text(b, 3)
    assert out.read_text() == H.golden_filter_fa(3)
