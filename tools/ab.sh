#!/bin/bash
# A/B kernel timing of library variants on the synthetic corpora (development aid; run under gpurun).
#   tools/ab.sh [variant.so ...]   -- the in-tree library is always timed last
#   AB_WL="tweets 1000000 4 3;docs 3000 4 3" tools/ab.sh ...   -- choose the workloads
IFS=';' read -ra WLS <<< "${AB_WL:-tweets 1000000 4 3;mixed 1000000 4 3;mixed 1000000 4 7;docs 3000 4 3}"
for lib in "$@" ""; do
  for wl in "${WLS[@]}"; do
    echo "== ${lib:-in-tree} $wl"
    LATOK_B200_LIB=$lib python tools/prof_run.py $wl 2>&1 | tail -1
  done
done
