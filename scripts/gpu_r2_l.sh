#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for b in 32 1; do
for g in 96 24 32 48 64 128 148; do
  env VTQ_DIFFNET_G=$g timeout 120 python scripts/diffnet_time.py $b 2>&1 | tail -1
done; done
