#!/bin/bash
# Round 2, run J: full GPU suite with the new attention kernel + cfg2 / cfg4 bench lines (new vs round-1 kernel)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -5
for cfg in cfg2 cfg4; do
for v in "" "VTQ_ATTN_V3=1"; do
  env $v timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('[$cfg $v] value',d['value'],'ms',d['ms_per_step'],'attn',k['attention']['avg_ms'],'frac',k['attention']['frac_of_burst'],'sustained',d['sustained'] and d['sustained']['value'],'clk',d['clocks']['sm_mhz'])"
done; done
