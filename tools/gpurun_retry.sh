#!/bin/bash
# dev helper: keep asking for a GPU slot until the call runs (exit code 3 = pod busy, nothing charged)
log=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > "$log" 2>&1; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
