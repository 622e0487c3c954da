#!/bin/bash
# Micro-benchmarks behind the design decisions (build first: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
# -I../../vtamiq_b200/csrc -o <name> <name>.cu in scripts/ubench/): tensor-pipe chains, SFU body, row-max pass.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for b in mma_chain sfu_pipe fmnmx; do
  timeout 200 scripts/ubench/$b > gpurun_out/ubench_$b.txt 2>&1; echo "=== $b rc=$?"; cat gpurun_out/ubench_$b.txt
done
