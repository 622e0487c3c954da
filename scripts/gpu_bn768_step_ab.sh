#!/bin/bash
# Same-box A/B of the N = 768 tile-width rule on whole steps (graph replay): forced 192 vs the M-aware default.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
out=gpurun_out/bn768_step_ab.txt; : > $out
for rep in 1 2 3; do
  for cfg in cfg2 cfg3 cfg5; do
    for bn in 192 auto; do
      if [ $bn = auto ]; then unset VTQ_GEMM_BN_N768; else export VTQ_GEMM_BN_N768=$bn; fi
      python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu --no-sustained 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('$cfg bn=$bn rep=$rep', d['value'], 'pairs/s', d['ms_per_step'], 'ms', 'out', k['gemm_out']['avg_ms'], 'fc2', k['gemm_fc2']['avg_ms'], 'clk', d['clocks']['sm_mhz'])" >> $out
    done
  done
done
cat $out
