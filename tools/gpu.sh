#!/bin/bash
# retry wrapper around gpurun: tools/gpu.sh <timeout_s> <logname> '<command>'   (retries while the pod answers busy / draining)
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "${@:1:$(($#-1))}" -- "${@: -1}" > gpurun_out/$LOG 2>&1
  rc=$?
  if grep -q "status=transient\|exit code 3\|status=busy" gpurun_out/$LOG || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
exit $rc
