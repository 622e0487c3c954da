#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for v in "" wait1 wait2; do
  env VTQ_LIBRARY=${v:+$PWD/vtamiq_b200/variants/lib_$v.so} timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "attention" 2>&1 | tail -1 | sed "s/^/[$v] /"
  env VTQ_LIBRARY=${v:+$PWD/vtamiq_b200/variants/lib_$v.so} timeout 120 python scripts/attn_time.py 2>&1 | tail -1 | sed "s/^/[$v] /"
  env VTQ_LIBRARY=${v:+$PWD/vtamiq_b200/variants/lib_$v.so} timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('[$v] value',d['value'],'ms',d['ms_per_step'],'attn',k['attention']['avg_ms'],'qkv',k['gemm_qkv']['avg_ms'],'fc1',k['gemm_fc1']['avg_ms'],'fc2',k['gemm_fc2']['avg_ms'],'out',k['gemm_out']['avg_ms'],'clk',d['clocks']['sm_mhz'])"
done
