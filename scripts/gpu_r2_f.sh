#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
L=$PWD/vtamiq_b200/variants/lib_noprescan.so
env VTQ_LIBRARY=$L timeout 120 python scripts/attn_trace.py > gpurun_out/attn_trace_v5_noprescan.txt 2>&1; head -28 gpurun_out/attn_trace_v5_noprescan.txt | cut -c1-200
env VTQ_ATTN_TURNS=1 VTQ_LIBRARY=$L timeout 120 python scripts/attn_trace.py > gpurun_out/attn_trace_v5_noprescan_turns.txt 2>&1; head -28 gpurun_out/attn_trace_v5_noprescan_turns.txt | cut -c1-200
