#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer run of the two native host parsers (LZ4 frame decoder, .dot text
# parser) over corpora of valid and randomly damaged inputs, with exact-size heap buffers so that any over-read or
# over-write is reported.  Run from the repo root:   bash tools/sanitize/run.sh
# Last run (round 1): 1 938 LZ4 files and 745 .dot files, two thread settings each -- no report.
set -e
here=$(cd "$(dirname "$0")" && pwd); root=$(cd "$here/../.." && pwd); work=$(mktemp -d)
flags="-O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer"
g++ $flags "$here/lz4_harness.cpp" "$root/nfact_b200/csrc/lz4_frame.cpp" -o "$work/lz4_harness" -lpthread
g++ $flags "$here/dot_harness.cpp" "$root/nfact_b200/csrc/dot_parser.cpp" -o "$work/dot_harness" -lpthread
PYTHONPATH="$root" python "$here/make_corpus.py" "$work"
ASAN_OPTIONS=detect_leaks=0 "$work/lz4_harness" "$work"/lz4/*
# the same files with linked frames split into side-by-side chunks (scratch streams + taint tracking)
NFB_LZ4_CHUNK_BYTES=131072 NFB_LZ4_WARM_BYTES=65536 ASAN_OPTIONS=detect_leaks=0 "$work/lz4_harness" "$work"/lz4/*
ASAN_OPTIONS=detect_leaks=0 "$work/dot_harness" "$work"/dot/*
rm -rf "$work"
