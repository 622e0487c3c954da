#!/bin/bash
# BASELINE configs[4]: throughput vs batch (pairs per step) on one B200; results -> gpurun_out/sweep.jsonl
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
for b in 32 64 128 256 512 1024; do
  timeout 600 python bench.py --pairs $b --steps 10 --warmup 3 --no-cpu 2>/dev/null | tail -1 >> gpurun_out/sweep.jsonl
  python - <<PY
import json
l=open("gpurun_out/sweep.jsonl").read().strip().splitlines()[-1]
d=json.loads(l); print("pairs",d["config"]["pairs_per_gpu"],"value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"gemm",d["roofline"]["achieved"])
PY
done
