#!/bin/sh
# GPU drop-in with the calling convention of Phylign's run_cobs_streaming.sh: five positionals in,
# `cobs query` text on stdout.  The xz stream is decoded on the host, the query runs on the B200
# (python -m phylign_b200.cli run-cobs-streaming); there is no CPU fallback.
[ "$#" -eq 5 ] || {
	echo "usage: $(basename -- "$0") kmer_thres threads cobs_index.xz uncompressed_size query.fa" >&2
	exit 1
}
repo=$(CDPATH= cd -- "$(dirname -- "$0")/.." && pwd) || exit 1
PYTHONPATH="$repo${PYTHONPATH:+:$PYTHONPATH}" exec python3 -m phylign_b200.cli run-cobs-streaming "$@"
