#!/bin/bash
# each attention case in its own process (a fault poisons the context)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for k in "2-129-12-fp16" "4-501-12-fp16" "2-128-2-fp16" "3-65-12-fp16" "1-1-12-fp16" "2-257-12-fp16" "1-1300-4-bf16"; do
  timeout 120 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "test_attention and $k" 2>&1 | tail -4 | grep -E "passed|failed|Error|error" | head -3 | sed "s/^/[$k] /"
done
