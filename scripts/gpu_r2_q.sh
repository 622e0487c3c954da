#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
VTQ_SEQ_PARTS=2 timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu -p no:cacheprovider -x -k "golden or parity_batch32" 2>&1 | tail -2
for rep in 1 2; do
for cfg in cfg2 cfg5; do
for v in 1 2 4; do
  env VTQ_SEQ_PARTS=$v timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$cfg parts=$v] value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'sustained',d['sustained'] and d['sustained']['value'],'clk',d['clocks']['sm_mhz'])"
done; done; done
