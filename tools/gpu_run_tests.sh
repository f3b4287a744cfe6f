#!/bin/bash
# Runs each GPU test file in its own process (a device trap in one file cannot poison the rest)
# and keeps the logs under gpurun_out/.
mkdir -p gpurun_out
rc=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -x -q -s --timeout 600 > "gpurun_out/${name}.log" 2>&1
  r=$?
  echo "== $f rc=$r"
  tail -n 40 "gpurun_out/${name}.log"
  [ $r -ne 0 ] && rc=$r
done
exit $rc
