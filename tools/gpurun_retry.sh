#!/bin/bash
# usage: gpurun_retry.sh <logfile> <gpurun args...>   -- resubmits while the pod answers "transient" (nothing charged), up to 20 times
log="$1"; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient" "$log"; then exit 0; fi
  sleep 45
done
