#!/bin/bash
# run tools/perf_probe.py once per tuning build under build/variants/ (kernel-time comparison on one GPU box)
mb=${1:-16}
for f in build/variants/*.so; do
  n=$(basename $f .so)
  GROMGPU_LIB=$PWD/$f python tools/perf_probe.py --mb $mb --reps 3 --skip-e2e 1 --skip-cnv 1 2>&1 | grep -E "pileup:|Error|error" | tail -1 | sed "s/^/$n: /"
done
