#!/bin/bash
# usage: tools/gpu_retry.sh <tag> <timeout_s> '<command>'   (retries while the pod answers busy; log in gpurun_out/<tag>.call)
tag=$1; to=$2; cmd=$3
cd /root/repo
mkdir -p gpurun_out
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$cmd" > gpurun_out/$tag.call 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/$tag.call; then echo "done rc=$rc" >> gpurun_out/$tag.call; exit $rc; fi
  sleep 90
done
echo "gave up" >> gpurun_out/$tag.call
