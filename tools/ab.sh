#!/bin/bash
# usage: tools/ab.sh "<nvcc extra flags>" "<env assignments>" [bench args...]  -- rebuild with flags, run bench, print one summary line
extra="$1"; envs="$2"; shift 2
LIODOM_NVCC_EXTRA="$extra" python -c "from liodom_b200 import build; build.build(force=True)" >/dev/null 2>&1 || echo BUILD FAILED
echo "== flags[$extra] env[$envs] args[$*]"
env $envs timeout 120 python bench.py --no-cpu-baseline "$@" | python tools/benchline.py
