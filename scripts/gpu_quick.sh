#!/bin/bash
# quick loop: selected GPU tests + bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -k "${1:-test_}" 2>&1 | tail -25 > gpurun_out/pytest_quick.log
echo "=== pytest: $(tail -1 gpurun_out/pytest_quick.log)"; grep -E "FAILED|^E  " gpurun_out/pytest_quick.log | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "=== bench rc=$?"; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_quick.json"))
    print("value",d["value"],"ms/step",d["ms_per_step"],"e2e",d["e2e"]["value"],"gemm",d["roofline"]["achieved"],d["roofline"]["frac"])
    for k,v in d["kernels"].items(): print(f"  {k:16s} n={v['launches_per_step']:3d} avg_ms={v['avg_ms']:.4f} share={v['share_of_step']:.3f} tflops={v.get('tflops','')}")
except Exception as e:
    print("bench parse failed",e); print(open("gpurun_out/bench_quick.err").read()[-2000:])
PY
