#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm2_kernel" -s 60 -c 4 -o gpurun_out/prof_gemm2_b -f \
     python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-sustained > gpurun_out/ncu_prof_gemm2_b.log 2>&1
echo "rc=$?"
