#!/usr/bin/env python3
"""Harness self-test only: the oracle CLI posing as `cobs` (classic-construct DIR OUT / query ...),
so that tests/test_real_cobs.py can be proven to run end to end without the real binary."""
import os, subprocess, sys, tempfile
ORC = os.environ["ORC_CLI"]
a = sys.argv[1:]
if a[0] == "classic-construct":
    a = [x for x in a[1:] if x != "--clobber"]
    ddir, out = a[-2], a[-1]
    opts = a[:-2]
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        for fn in sorted(os.listdir(ddir)):
            f.write(open(os.path.join(ddir, fn)).read())
    sys.exit(subprocess.call([ORC, "construct", "-o", out] + opts + [f.name]))
sys.exit(subprocess.call([ORC] + a))
