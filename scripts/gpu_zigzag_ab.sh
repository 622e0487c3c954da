#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "golden or parity_batch32 or pruning or layernorm or gemm or attention or cfg3_full" 2>&1 | tail -2
for rep in 1 2 3; do
for cfg in cfg2 cfg5; do
for v in 1 0; do
  env VTQ_ZIGZAG=$v timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu --no-sustained 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('[$cfg zigzag=$v] value',d['value'],'ms',d['ms_per_step'],'ln',k['layernorm']['avg_ms'],'qkv',k['gemm_qkv']['avg_ms'],'attn',k['attention']['avg_ms'],'out',k['gemm_out']['avg_ms'],'fc1',k['gemm_fc1']['avg_ms'],'fc2',k['gemm_fc2']['avg_ms'],'clk',d['clocks']['sm_mhz'])"
done; done; done
