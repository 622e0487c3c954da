#!/bin/bash
# Round 2, run G: attention kernel v6 (4 softmax warpgroups + helper) — unit tests, stand-alone time vs v5 / v3, timeline.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "attention" 2>&1 | tail -15
for v in "" "VTQ_ATTN_V5=1" "VTQ_ATTN_V3=1"; do
  env $v timeout 120 python scripts/attn_time.py 2>&1 | tail -1
done
timeout 120 python scripts/attn_trace.py > gpurun_out/attn_trace_v6.txt 2>&1; grep -v "^MMA thread" gpurun_out/attn_trace_v6.txt | head -34
