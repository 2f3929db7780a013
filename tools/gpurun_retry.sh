#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-s> <gpus> '<command>'  — retries while the pod answers busy (rc 3)
T=$1; G=$2; CMD=$3
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $i answered busy; sleeping 90 s"
  sleep 90
done
exit 3
