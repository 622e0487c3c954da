#!/bin/bash
# per-kernel table of the bench under two values of an env toggle: gpu_ab_kernels.sh VAR v0 v1
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for v in "$2" "$3"; do
  env "$1=$v" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null > gpurun_out/bench_ab_$v.json
  python - "$1=$v" gpurun_out/bench_ab_$v.json <<'PY'
import json,sys
d=json.load(open(sys.argv[2]))
print(sys.argv[1],"value",d["value"],"ms/step",d["ms_per_step"],"e2e",d["e2e"]["value"],"clk",d["clocks"]["sm_mhz"])
tot=0
for k,v in sorted(d["kernels"].items()):
    t=v['launches_per_step']*v['avg_ms']; tot+=t
    print(f"  {k:16s} n={v['launches_per_step']:3d} avg_ms={v['avg_ms']:.4f} total={t:.3f} tflops={v.get('tflops','')}")
print("  sum of kernels",round(tot,3))
PY
done
