#!/bin/bash
# A/B an environment toggle on the bench: usage gpu_ab.sh VAR val0 val1
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for v in "$2" "$3" "$2" "$3"; do
  env "$1=$v" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1=$v value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clk',d['clocks']['sm_mhz'])"
done
