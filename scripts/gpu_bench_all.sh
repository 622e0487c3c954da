#!/bin/bash
# Bench lines for every BASELINE config + same-shape cuBLAS GEMMs next to vtq_gemm.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
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
timeout 300 python scripts/gemm_vs_cublas.py 32 > gpurun_out/gemm_vs_cublas.jsonl 2> gpurun_out/gemm_vs_cublas.err; echo "=== cublas rc=$?"; cat gpurun_out/gemm_vs_cublas.jsonl
timeout 300 python scripts/gemm_vs_cublas.py 256 >> gpurun_out/gemm_vs_cublas.jsonl 2>> gpurun_out/gemm_vs_cublas.err; tail -4 gpurun_out/gemm_vs_cublas.jsonl
