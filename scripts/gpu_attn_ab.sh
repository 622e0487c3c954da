#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "attention" 2>&1 | tail -1
for rep in 1 2 3; do for v in "" before; do
  env VTQ_LIBRARY=${v:+$PWD/vtamiq_b200/variants/lib_$v.so} timeout 120 python scripts/attn_time.py 2>&1 | tail -1 | sed "s/^/[$v] /"
done; done
timeout 120 python scripts/attn_trace.py 2>&1 | grep -v "^MMA" | head -4
