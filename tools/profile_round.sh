#!/bin/bash
# One GPU call's worth of profiling evidence (run under gpurun from the repo root):
#   tools/profile_round.sh <tag> [lanes]
# 1. launch list of two steady-state steps (ncu gpu__time_duration.sum, --clock-control none)
# 2. ncu --set full of every kernel of one steady-state step (lane groups forced to 1: one launch per stage)
# 3. ncu --set full --import-source on of the association kernel alone (source-level stalls)
# Outputs land in gpurun_out/ (scratch); tools/ncu_summary.py turns them into profiles/*.txt here.
TAG=${1:-r02}
LANES=${2:-128}
export LIODOM_LANE_GROUPS=1
PER=18                       # launches per step with one lane group (incremental hash: 4 build kernels)
PRE=$(( (16 + 3) * PER ))    # pre-roll 16 + warm-up 3 steps
BENCH="python bench.py --lanes $LANES --steps 2 --warmup 3 --no-cpu-baseline --no-stage-pass --no-single-stream --no-sharded-block --no-xyz12"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s $PRE -c $(( 2 * PER )) --csv --log-file gpurun_out/launches_${TAG}_lanes${LANES}.csv $BENCH > gpurun_out/prof_${TAG}_a.log 2>&1
ncu --set full --clock-control none -s $PRE -c $PER -o gpurun_out/step_${TAG}_lanes${LANES} -f $BENCH > gpurun_out/prof_${TAG}_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_associate -s $(( (16 + 3) * 2 )) -c 2 -o gpurun_out/k_associate_${TAG}_lanes${LANES} -f $BENCH > gpurun_out/prof_${TAG}_c.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_${TAG}_lanes${LANES}.csv
