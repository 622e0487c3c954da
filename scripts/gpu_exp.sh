#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for v in 0 1; do
  VTQ_ATT_DBG=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('dbg=$v value',d['value'],'attn',d['kernels']['attention'])"
done
