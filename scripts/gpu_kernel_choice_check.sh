#!/bin/bash
# Check of the final kernel-choice rule: rule vs pair at 1 / 2 / 4 pairs, GEMM + forward tests, smoke, cfg1 + cfg2 lines.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out/final_r2
out=gpurun_out/kernel_choice_check.txt; : > $out
run () {
  local label=$1; shift
  python bench.py "$@" --steps 100 --warmup 10 --no-cpu --no-sustained 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('$label', d['value'], 'pairs/s', d['ms_per_step'], 'ms', 'qkv', k['gemm_qkv']['avg_ms'], 'out', k['gemm_out']['avg_ms'], 'fc1', k['gemm_fc1']['avg_ms'], 'fc2', k['gemm_fc2']['avg_ms'])" >> $out
}
for spec in "--config cfg2 --pairs 2" "--config cfg2 --pairs 4" "--config cfg2 --pairs 3"; do
  unset VTQ_GEMM_1CTA;    run "rule  [$spec]" $spec
  export VTQ_GEMM_1CTA=0; run "pair  [$spec]" $spec
done
unset VTQ_GEMM_1CTA
cat $out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py -q -m gpu -p no:cacheprovider -x -k "gemm or golden or forward_matches or variants or batch" 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 600 python bench.py --config cfg1 --steps 20 --warmup 5 > gpurun_out/final_r2/bench_cfg1.json 2>/dev/null; echo "cfg1 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/final_r2/bench_cfg1.json')); print('cfg1', d['value'], d['ms_per_step'], d['e2e']['value'])"
