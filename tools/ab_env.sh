#!/bin/bash
# usage: tools/ab_env.sh "<env assignments>" [bench args...] -- run the bench (no CPU legs) under an environment, print one summary line
envs="$1"; shift
echo "== env[$envs] args[$*]"
env $envs timeout 300 python bench.py --no-cpu-baseline --no-single-stream "$@" | python tools/benchline.py
