#!/bin/bash
# keep asking for a GPU slot until the call actually runs (exit code 3 / "transient" = nothing charged)
# usage: scripts/gpu/retry.sh <timeout_s> <out_file> <command...>
T=$1; OUT=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $OUT 2>&1
  if grep -q "status=transient\|retry in a few minutes\|rc=None" $OUT; then sleep 90; continue; fi
  break
done
