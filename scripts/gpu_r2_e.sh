#!/bin/bash
# Round 2, run E: why are concurrent exponential phases slow?  De-phased SFU micro-benchmark + kernel variants.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 200 scripts/ubench/sfu_pipe > gpurun_out/sfu_pipe_skew.txt 2>&1; grep -E "^1 |^0 |de-phased" gpurun_out/sfu_pipe_skew.txt
for v in "" noprescan sync sync_noprescan; do
  env VTQ_LIBRARY=${v:+$PWD/vtamiq_b200/variants/lib_$v.so} timeout 120 python scripts/attn_time.py 2>&1 | tail -1 | sed "s/^/[$v] /"
  env VTQ_ATTN_TURNS=1 VTQ_LIBRARY=${v:+$PWD/vtamiq_b200/variants/lib_$v.so} timeout 120 python scripts/attn_time.py 2>&1 | tail -1 | sed "s/^/[$v] /"
done
env ATT_FINE=1 VTQ_LIBRARY=$PWD/vtamiq_b200/variants/lib_fine.so timeout 120 python scripts/attn_trace.py > gpurun_out/attn_trace_v5_fine.txt 2>&1; head -30 gpurun_out/attn_trace_v5_fine.txt
