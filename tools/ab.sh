#!/bin/bash
# A/B timing of two engine builds on one GPU, interleaved: tools/ab.sh <workload> <rounds> [tune]
# A = tools/ab/libh2sha_base.so (build of an earlier commit), B = the in-tree library.
wl=${1:-cfg2}; n=${2:-3}; tune=${3:-"parts=3"}
for i in $(seq $n); do
  echo -n "A "; TUNE_SUSTAIN=1 TUNE_LIB=tools/ab/libh2sha_base.so python tools/tune.py $wl "$tune" | tail -1
  echo -n "B "; TUNE_SUSTAIN=1 python tools/tune.py $wl "$tune" | tail -1
done
