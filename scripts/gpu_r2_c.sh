#!/bin/bash
# Round 2, run C: tensor-pipe chain micro-benchmark + attention timeline of the shipped kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 120 scripts/ubench/mma_chain > gpurun_out/mma_chain.txt 2>&1; echo "=== mma_chain rc=$?"; cat gpurun_out/mma_chain.txt
timeout 120 python scripts/attn_time.py 2>&1 | tail -1
