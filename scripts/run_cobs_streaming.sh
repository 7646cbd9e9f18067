#!/usr/bin/env bash
# Drop-in for the reference's scripts/run_cobs_streaming.sh (same 5 positionals, same stdout):
# the index is decompressed on the host, queried on the B200.  No CPU fallback.
set -e
set -o pipefail
set -u

readonly PROGNAME=$(basename "$0")
readonly REPO=$(cd "$(dirname "$0")/.." && pwd)
if [[ $# -ne 5 ]]; then
	>&2 echo "usage: $PROGNAME kmer_thres threads cobs_index.xz uncompressed_size query.fa"
	exit 1
fi
PYTHONPATH="${REPO}${PYTHONPATH:+:$PYTHONPATH}" exec python3 -m phylign_b200.cli run-cobs-streaming "$1" "$2" "$3" "$4" "$5"
