#!/bin/bash
# usage: tools/gpurun_retry.sh <out-file> <gpurun args...>   -- retries while the pod answers "no slot" (exit 3)
out=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)" >> "$out"; exit $rc; fi
  sleep 90
done
echo "gpurun: gave up after 40 attempts" >> "$out"; exit 3
