#!/bin/bash
# A/B an env toggle: kernel tests under each value, then alternating bench runs.  usage: gpu_ab_tests.sh VAR v0 v1 "<-k expr>"
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for v in "$2" "$3"; do
  env "$1=$v" timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "${4:-attention}" 2>&1 | tail -12 > gpurun_out/pytest_ab_$v.log
  echo "=== $1=$v pytest: $(tail -1 gpurun_out/pytest_ab_$v.log)"; grep -E "FAILED|^E  .*assert" gpurun_out/pytest_ab_$v.log | head -8
done
for v in "$2" "$3" "$2" "$3" "$2" "$3"; do
  env "$1=$v" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1=$v value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'attn',d['kernels']['attention']['avg_ms'],'clk',d['clocks']['sm_mhz'])"
done
