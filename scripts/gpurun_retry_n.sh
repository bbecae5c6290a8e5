#!/usr/bin/env bash
# like gpurun_retry.sh with --gpus N.  Usage: scripts/gpurun_retry_n.sh <gpus> <timeout_s> <logfile> <command...>
G="$1"; T="$2"; LOG="$3"; shift 3
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 8
done
exit 3
