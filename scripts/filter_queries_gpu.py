#!/usr/bin/env python3
"""Same CLI as the reference's scripts/filter_queries.py (-n N -q query.fa match files...),
merge done on the GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phylign_b200.cli import main  # noqa: E402

main(["filter"] + sys.argv[1:])
