#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for v in "" skipA skipB; do
  env VTQ_LIBRARY=${v:+$PWD/vtamiq_b200/variants/lib_$v.so} timeout 300 python scripts/gemm_vs_cublas.py 32 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('[$v]', d['shape'], d['vtq_gemm']['ms'], d['vtq_gemm']['tflops'])"
done
