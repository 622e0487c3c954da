#!/bin/bash
# A/B several builds of the library (same ABI): gpu_ab_libs.sh <lib.so> [<lib.so> ...]; "" = the in-tree build
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for rep in 1 2 3; do
  for v in "" "$@"; do
    env VTQ_LIBRARY=${v:+$PWD/$v} timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lib=${v:-default} value',d['value'],'ms',d['ms_per_step'],'attn',d['kernels']['attention']['avg_ms'],'clk',d['clocks']['sm_mhz'])"
  done
done
