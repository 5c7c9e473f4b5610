#!/bin/bash
# usage: tools/gpu_retry.sh [gpurun options ...] -- '<command>'
# Retries a gpurun call while the pod answers "busy / draining" (exit code 3, nothing charged).
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpu_retry] busy (attempt $attempt), sleeping 150 s" >&2
  sleep 150
done
exit 3
