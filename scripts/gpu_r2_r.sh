#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:diffnet_fused" -s 1 -c 1 -o gpurun_out/prof_diffnet -f \
     python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-sustained > gpurun_out/ncu_prof_diffnet.log 2>&1
echo "rc=$?"
