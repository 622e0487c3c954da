#!/bin/bash
# 1-CTA vs CTA-pair GEMM kernel at small batches: the shipped rule (unset) vs either kernel forced; then the GPU tests.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
out=gpurun_out/kernel_choice_ab.txt; : > $out
run () {  # label, bench args..., env through VTQ_GEMM_1CTA_SET
  local label=$1; shift
  python bench.py "$@" --steps 100 --warmup 10 --no-cpu --no-sustained 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('$label', d['value'], 'pairs/s', d['ms_per_step'], 'ms', 'qkv', k['gemm_qkv']['avg_ms'], 'out', k['gemm_out']['avg_ms'], 'fc1', k['gemm_fc1']['avg_ms'], 'fc2', k['gemm_fc2']['avg_ms'], 'clk', d['clocks']['sm_mhz'])" >> $out
}
for spec in "--config cfg1" "--config cfg2 --pairs 1" "--config cfg2 --pairs 2" "--config cfg2 --pairs 4" "--config cfg2 --pairs 8"; do
  unset VTQ_GEMM_1CTA;    run "rule  [$spec]" $spec
  export VTQ_GEMM_1CTA=0; run "pair  [$spec]" $spec
  export VTQ_GEMM_1CTA=1; run "1cta  [$spec]" $spec
done
unset VTQ_GEMM_1CTA
cat $out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
