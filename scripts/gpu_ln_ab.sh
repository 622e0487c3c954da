#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "layernorm" 2>&1 | tail -1
for rep in 1 2 3; do
for v in 1 0; do
  env VTQ_LN_REVERSE=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-sustained 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('[ln reverse $v] value',d['value'],'ms',d['ms_per_step'],'layernorm',k['layernorm']['avg_ms'],'qkv',k['gemm_qkv']['avg_ms'],'fc1',k['gemm_fc1']['avg_ms'],'clk',d['clocks']['sm_mhz'])"
done; done
