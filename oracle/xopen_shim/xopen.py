"""Test-only stand-in for the `xopen` package (0.7.3 in the reference's conda env, absent here):
the only third-party import of /root/reference/scripts/filter_queries.py:10,13,40.  Put on PYTHONPATH
when the unmodified script is executed by the golden-vector generator and by bench.py's reference arm."""
import gzip


def xopen(fn, mode="r"):
    fn = str(fn)
    return gzip.open(fn, mode + "t") if fn.endswith(".gz") else open(fn, mode)
