#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "gemm" 2>&1 | tail -3
for rep in 1 2; do
for v in "" nbuf1; do
  env VTQ_LIBRARY=${v:+$PWD/vtamiq_b200/variants/lib_$v.so} timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-sustained 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('[$v] value',d['value'],'ms',d['ms_per_step'],'out',k['gemm_out']['avg_ms'],'fc2',k['gemm_fc2']['avg_ms'],'embed',k['gemm_embed']['avg_ms'],'qkv',k['gemm_qkv']['avg_ms'],'clk',d['clocks']['sm_mhz'])"
done; done
timeout 300 python scripts/gemm_vs_cublas.py 32 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['shape'], d['vtq_gemm'], d['cublas_matmul_nt'])"
