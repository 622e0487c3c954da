#!/bin/bash
# A/B two builds of the library (same ABI): vtamiq_b200/libvtamiq_b200_prev.so vs the in-tree build.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
PREV=$PWD/vtamiq_b200/libvtamiq_b200_prev.so
for v in "$PREV" "" "$PREV" "" "$PREV" ""; do
  env VTQ_LIBRARY=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lib=${v:+prev} value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'attn',d['kernels']['attention']['avg_ms'],'clk',d['clocks']['sm_mhz'])"
done
