#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / status=transient): tools/gpurun_retry.sh <timeout> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then OUT=$(gpurun --timeout $T -- "$CMD" 2>&1); else OUT=$(gpurun --gpus $G --timeout $T -- "$CMD" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up: pod busy"; exit 3
