#!/usr/bin/env python3
"""Record the DRAM traffic of one `ncu --set full` capture in profiles/traffic.json, keyed by workload
AND by the sha256 of the kernel sources it was taken from (bench.py ignores entries whose hash differs
from the sources it runs, so a stale capture can never label a newer kernel).

    python scripts/update_traffic.py REPORT.ncu-rep reads1k 100000 1000 64 4000 1 prune1
"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

rep, name, n_reads, read_len, n_idx, n_docs, world, prune = sys.argv[1:9]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]


def metric(m):
    v, u = float(vals[hdr.index(m)].replace(",", "")), units[hdr.index(m)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


kernel = "gather_count_ring_kernel<32,10,3,4>"
key = f"{kernel}|{name}|{n_reads}|{read_len}|{n_idx}|{n_docs}|{world}|{prune}"
p = os.path.join(ROOT, "profiles", "traffic.json")
tj = json.load(open(p)) if os.path.exists(p) else {}
tj[key] = {"traffic_bytes": int(metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")),
           "dram_bytes_read": int(metric("dram__bytes_read.sum")), "dram_bytes_write": int(metric("dram__bytes_write.sum")),
           "kernel_name": vals[hdr.index("Kernel Name")], "duration_ms": float(vals[hdr.index("gpu__time_duration.sum")].replace(",", "")),
           "kernel_source_sha256": bench.kernel_source_digest(), "source": os.path.basename(rep)}
json.dump(tj, open(p, "w"), indent=1)
print(key, tj[key])
