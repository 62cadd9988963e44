#!/bin/bash
# each GPU test file on its own, bounded, verbose progress into gpurun_out/t_<file>.log
for f in "$@"; do
  n=$(basename $f .py)
  timeout -k 5 150 python -u -m pytest $f -m gpu -x -v --timeout 120 -p no:cacheprovider > gpurun_out/t_$n.log 2>&1
  echo "$n rc=$? $(tail -1 gpurun_out/t_$n.log)"
done
