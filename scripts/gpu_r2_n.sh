#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "gemm" 2>&1 | tail -2
for rep in 1 2 3; do
for v in "" before; do
  env VTQ_LIBRARY=${v:+$PWD/vtamiq_b200/variants/lib_$v.so} timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-sustained 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('[$v] value',d['value'],'ms',d['ms_per_step'],'out',k['gemm_out']['avg_ms'],'fc2',k['gemm_fc2']['avg_ms'],'fc1',k['gemm_fc1']['avg_ms'],'qkv',k['gemm_qkv']['avg_ms'],'clk',d['clocks']['sm_mhz'])"
done; done
