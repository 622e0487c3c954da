#!/bin/bash
# Latency configuration (cfg1: 1 pair, 256 patches): host-side choices A/B on one box.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
out=gpurun_out/cfg1_ab.txt; : > $out
run () {  # label, env...
  local label=$1; shift
  env "$@" python bench.py --config cfg1 --steps 200 --warmup 20 --no-cpu --no-sustained 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('$label', d['value'], 'pairs/s', d['ms_per_step'], 'ms', 'diffnet', k['diffnet_head']['avg_ms'], 'qkv', k['gemm_qkv']['avg_ms'], 'fc2', k['gemm_fc2']['avg_ms'], 'clk', d['clocks']['sm_mhz'])" >> $out
}
for rep in 1 2; do
  run base X=1
  run gemm_1cta VTQ_GEMM_1CTA=1
  run zigzag_off VTQ_ZIGZAG=0
  run diffnet_g64 VTQ_DIFFNET_G=64
  run diffnet_g96 VTQ_DIFFNET_G=96
  run diffnet_g148 VTQ_DIFFNET_G=148
  run bn768_256 VTQ_GEMM_BN_N768=256
  run bn768_128 VTQ_GEMM_BN_N768=128
done
cat $out
