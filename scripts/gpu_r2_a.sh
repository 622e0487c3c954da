#!/bin/bash
# Round 2, run A: GPU test suite with the new tests, smoke, bench lines for every BASELINE config (baseline of the round).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt; nvidia-smi -L >> gpurun_out/host.txt
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider -s -x --durations=15 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
echo "=== pytest: $(grep -E 'passed|failed|error' gpurun_out/pytest_gpu.log | tail -1)"; grep -E "^N=|^B=|cfg4|tail gradients|FAILED|^E  |Error" gpurun_out/pytest_gpu.log | head -40
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "=== smoke rc=$?"; tail -2 gpurun_out/smoke.log
for cfg in cfg2 cfg1 cfg3 cfg4 cfg5; do
  timeout 900 python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  echo "=== bench $cfg rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$cfg.json").read())
    r=d["roofline"]; s=d["sustained"]
    print(d["value"], "pairs/s", d["ms_per_step"], "ms/step e2e", d["e2e"]["value"], "e2e_img", d["e2e_from_images"]["value"], "gemm", r["achieved"], r["frac_of_burst"], "algo frac burst", d["frac_of_bf16_peak"]["burst"], "sustained", s and s["value"], s and s["gemm"]["frac"], "cpu", d["cpu_baseline"])
    for k,v in d["kernels"].items(): print("   ", k, v)
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_$cfg.err").read()[-1500:])
PY
done
