#!/bin/bash
# Round 2, run K: cluster DiffNet kernel — tests (kernel + end-to-end goldens), bench A/B against the cooperative kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "diffnet or golden or tail or smoke or cfg5 or parity_batch32" 2>&1 | tail -4
for cfg in cfg2 cfg1 cfg5; do
for v in "" "VTQ_DIFFNET_COOP=1"; do
  env $v timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('[$cfg $v] value',d['value'],'ms',d['ms_per_step'],'diffnet',k['diffnet_head']['avg_ms'],'clk',d['clocks']['sm_mhz'])"
done; done
