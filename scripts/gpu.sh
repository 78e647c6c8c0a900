#!/bin/bash
# Runs a command on the GPU box through gpurun, retrying while the pod answers "busy / transient" (nothing is charged
# for those).  usage: scripts/gpu.sh [--gpus N] TIMEOUT_S 'command'
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@"
  rc=$?
  st=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json'))['status'])" 2>/dev/null)
  if [ "$st" != "transient" ] && [ "$rc" != "3" ]; then exit $rc; fi
  echo "[gpu.sh] attempt $attempt: pod busy, retrying in 90 s"
  sleep 90
done
exit 3
