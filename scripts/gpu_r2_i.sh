#!/bin/bash
# ncu --set full of the attention kernel alone (cfg2 shape, stand-alone launches), source-level stall samples
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:attention_kernel" -s 6 -c 1 -o gpurun_out/prof_attention_v6 -f python scripts/attn_time.py > gpurun_out/ncu_attention_v6.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_attention_v6.log; ls -la gpurun_out/*.ncu-rep
