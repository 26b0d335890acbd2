#!/bin/bash
# usage: scripts/gpurun_retry.sh [gpurun flags...] -- 'command'   — retries while the pod answers busy / transient
for attempt in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient"; then echo "[retry $attempt] pod busy, sleeping 150 s"; sleep 150; continue; fi
  exit $rc
done
exit 3
