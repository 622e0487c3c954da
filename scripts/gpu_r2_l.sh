#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for v in 0 1 2 4 8 5 3 10 15 7; do
  env VTQ_DC_DBG=$v timeout 120 python scripts/diffnet_time.py 32 2>&1 | tail -2 | head -1
done
