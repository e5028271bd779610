#!/bin/bash
# compute-sanitizer memcheck + racecheck of every step layout on a small mesh with the limiter tripping
# (remediation path) -- and, with N >= 2 GPUs, of the overlapped two-stream 2-rank step through the
# C++ drop-in.  Logs go to gpurun_out/sanitizer_*.log; summary lines at the end of each.
#   bash tools/sanitize.sh [ngpus]
set -u
cd "$(dirname "$0")/.."
NG=${1:-1}
OUT=gpurun_out
mkdir -p $OUT
W=$(mktemp -d)
python - "$W" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from cfd2d_b200 import cases
c = cases.channel(48, 24, jitter=0.2, shuffle=True, two_materials=True)
c.task.method = "FVM_TVD_CUDA"
c.task.p_max = 1.03e5
c.task.STEP_MAX = 6
c.task.FILE_OUTPUT_STEP = 3
c.task.LOG_OUTPUT_STEP = 3
c.write(sys.argv[1])
PY
BIN=$PWD/cfd-2d_b200/host/_build/cfd2d_cuda
CS=/usr/local/cuda/bin/compute-sanitizer
if [ -z "${SKIP1:-}" ]; then
for tool in memcheck racecheck; do
  for layout in 0 1 2; do
    for flux in GODUNOV LAX; do
      ( cd $W && sed "s#</task>#<gpu flux=\"$flux\" order=\"2\"/></task>#" task.xml > t_$flux.xml &&
        CFD2D_FUSED=$layout CFD2D_TILE=96 CFD2D_PIPE_TILE=64 timeout 600 $CS --tool $tool --error-exitcode 9 $BIN t_$flux.xml ) \
        > $OUT/sanitizer_${tool}_layout${layout}_${flux}.log 2>&1
      echo "exit $? $tool layout=$layout flux=$flux: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitizer_${tool}_layout${layout}_${flux}.log | tail -1)"
    done
  done
done
fi
if [ -z "${SKIP1:-}" ]; then      # first-order Lax-Friedrichs: the single cell-parallel sweep k_cell_lf1
for tool in memcheck racecheck; do
  ( cd $W && sed "s#</task>#<gpu flux=\"LAX\" order=\"1\"/></task>#" task.xml > t_LAX1.xml &&
    timeout 600 $CS --tool $tool --error-exitcode 9 $BIN t_LAX1.xml ) > $OUT/sanitizer_${tool}_layout0_LAX_order1.log 2>&1
  echo "exit $? $tool layout=0 flux=LAX order=1: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitizer_${tool}_layout0_LAX_order1.log | tail -1)"
done
fi
if [ "$NG" -ge 2 ]; then
  for tool in memcheck racecheck; do
    for layout in 0 2; do
      ( cd $W && CFD2D_FUSED=$layout CFD2D_PIPE_TILE=64 timeout 900 python $OLDPWD/tools/launch_ranks.py 2 $CS --tool $tool --error-exitcode 9 $BIN task.xml ) \
        > $OUT/sanitizer_${tool}_2rank_layout${layout}.log 2>&1
      echo "exit $? $tool 2-rank layout=$layout: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitizer_${tool}_2rank_layout${layout}.log | tr '\n' ' ')"
    done
  done
fi
