#!/bin/bash
# scripts/gpurun_retry.sh [--gpus N] <timeout-seconds> <command...> -- runs gpurun, retrying while the pod answers "transient"
# (no box / slot free: nothing charged).  Build-container helper; not used on the GPU box.
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for attempt in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy"; then
    echo "[gpurun_retry] attempt $attempt: transient, sleeping 90 s" >&2
    sleep 90
    continue
  fi
  echo "$out"
  exit 0
done
echo "$out"
exit 3
