"""Resident match server: one process keeps COBS indexes in HBM across many `cobs query` calls.

The reference starts one `cobs query` process per batch and per query set, and pays the xz
decode + load every time (/root/reference/Snakefile:431-487).  With the indexes resident in
HBM a query set costs milliseconds, so the drop-in `cobs query --server SOCK ...` front end
only ships the request to this process.  Indexes are keyed by
(realpath, size, mtime) and evicted least-recently-used when HBM runs out.

    python -m phylign_b200.cli serve --socket /tmp/phylign.sock [--device 0] [--preload DIR --batches FILE]

Wire protocol (Unix stream socket): one JSON line request, one JSON line response header
{"ok": bool, "error": str, "len": n, ...} followed by n payload bytes.
"""
from __future__ import annotations

import json
import os
import socket
import socketserver
import sys
import time

from . import fasta
from .cobs_text import format_cobs_text_arrays


class MatchService:
    def __init__(self, device=0, hbm_budget=0):
        from .matcher import Matcher
        self.m = Matcher(device, hbm_budget)
        self.resident = {}     # key -> idx_id
        self.last_used = {}    # key -> monotonic time
        self.stats = {"queries": 0, "loads": 0, "hits": 0, "evictions": 0}

    @staticmethod
    def key_of(path):
        st = os.stat(path)
        return (os.path.realpath(path), st.st_size, int(st.st_mtime))

    def ensure_index(self, path):
        from ._lib import PhylignCudaError
        key = self.key_of(path)
        if key in self.resident:
            self.stats["hits"] += 1
        else:
            batch = os.path.basename(path).split(".cobs_classic")[0]
            while True:
                try:
                    self.resident[key] = self.m.load_index(path, batch=batch)
                    break
                except PhylignCudaError as e:
                    if "PHY_ERR_NOMEM" not in str(e) or not self.resident:
                        raise
                    victim = min(self.resident, key=lambda k: self.last_used.get(k, 0))
                    self.m.evict(self.resident.pop(victim))     # LRU, then retry
                    self.last_used.pop(victim, None)
                    self.stats["evictions"] += 1
            self.stats["loads"] += 1
        self.last_used[key] = time.monotonic()
        return self.resident[key]

    def query(self, req) -> bytes:
        """`cobs query` for one index; returns the stdout bytes."""
        idx = self.ensure_index(req["index"])
        hdr = self.m.indexes[idx].header
        if req.get("index_sizes") is not None and req["index_sizes"] != hdr.header_size + hdr.body_size:
            raise ValueError(f"--index-sizes {req['index_sizes']} != header {hdr.header_size} + body {hdr.body_size}")
        qf = fasta.QueryFile(req["query"])                  # flat arrays (native reader)
        out = []
        for q0, q1 in qf.block_ranges(int(req.get("query_block_bases", 2 * 10 ** 9))):
            self.m.set_queries_raw(qf.seqs, qf.soffs[q0:q1 + 1])
            # only this index takes part: the others stay resident but are not queried
            res = self.m.match(req["threshold"], top_n=req.get("top_n", 0), floor_mode=req.get("floor", False),
                               only=[idx])
            out.append(format_cobs_text_arrays(qf.headers, qf.hoffs[q0:q1 + 1], res, self.m.indexes[idx],
                                               strip_prefix=req.get("top_n", 0) > 0))
        self.stats["queries"] += 1
        return b"".join(out)

    def handle(self, req):
        cmd = req.get("cmd")
        if cmd == "query":
            return {}, self.query(req)
        if cmd == "status":
            return {"resident": [k[0] for k in self.resident], **self.stats}, b""
        if cmd == "preload":
            for p in req["paths"]:
                self.ensure_index(p)
            return {"resident": len(self.resident)}, b""
        if cmd == "shutdown":
            return {"bye": True}, b""
        raise ValueError(f"unknown command {cmd!r}")


class _Handler(socketserver.StreamRequestHandler):
    def handle(self):
        line = self.rfile.readline()
        if not line:
            return
        try:
            req = json.loads(line)
            extra, payload = self.server.service.handle(req)
            head = {"ok": True, "len": len(payload), **extra}
        except Exception as e:          # the client turns this into a non-zero exit
            req, payload = {}, b""
            head = {"ok": False, "error": f"{type(e).__name__}: {e}", "len": 0}
        self.wfile.write(json.dumps(head).encode() + b"\n" + payload)
        self.wfile.flush()
        if req.get("cmd") == "shutdown":
            self.server.stop = True


def serve(sock_path, device=0, hbm_budget=0, preload=()):
    if os.path.exists(sock_path):
        os.unlink(sock_path)
    svc = MatchService(device, hbm_budget)
    for p in preload:
        svc.ensure_index(p)
    with socketserver.UnixStreamServer(sock_path, _Handler) as srv:
        srv.service, srv.stop = svc, False
        print(f"[phylign_b200] serving on {sock_path} ({len(svc.resident)} indexes resident)", file=sys.stderr,
              flush=True)
        while not srv.stop:
            srv.handle_request()
    svc.m.close()
    if os.path.exists(sock_path):
        os.unlink(sock_path)


def request(sock_path, req, timeout=3600.0):
    """Client side: (header dict, payload bytes)."""
    with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as s:
        s.settimeout(timeout)
        s.connect(sock_path)
        s.sendall(json.dumps(req).encode() + b"\n")
        f = s.makefile("rb")
        head = json.loads(f.readline())
        payload = f.read(head.get("len", 0)) if head.get("len") else b""
    return head, payload
