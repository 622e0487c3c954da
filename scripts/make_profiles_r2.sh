#!/bin/bash
# Turns the scratch outputs of scripts/gpu_final_r2.sh (gpurun_out/final_r2/) into the tracked evidence under profiles/.
cd "$(dirname "$0")/.."; O=gpurun_out/final_r2; R=r02
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do cp $O/bench_$c.json profiles/${R}_bench_$c.json; done
cp $O/bench_cfg2.json profiles/${R}_bench.json
if [ -f $O/gemm_traffic.json ]; then cp $O/gemm_traffic.json profiles/gemm_traffic.json; cp $O/ncu_gemm2.txt profiles/${R}_ncu_gemm2.txt; fi
cp $O/bench_reference.json profiles/${R}_bench_reference.json
cp $O/launches.csv profiles/${R}_launches.csv
python scripts/launch_summary.py $O/launches.csv > profiles/${R}_launches.md
tail -2 $O/smoke.log > profiles/${R}_smoke.txt
cat $O/host.txt > profiles/${R}_host.txt
python - <<'PY'
import json
rows=[]
for c in ("cfg1","cfg2","cfg3","cfg4","cfg5"):
    d=json.load(open(f"profiles/r02_bench_{c}.json")); r=d["roofline"]; s=d["sustained"] or {}
    k=d["kernels"]
    rows.append(f"| {c} | {d['config']['pairs_per_step']} | {d['value']:.1f} | {d['ms_per_step']:.3f} | {d['e2e']['value']:.1f} | {d['e2e_from_images']['value']:.1f} | {s.get('value','')} | {r['achieved']:.0f} ({r['frac_of_burst']:.2f}) | {d['frac_of_bf16_peak']['burst']:.3f} | {k['attention']['avg_ms']:.4f} ({k['attention']['share_of_step']:.2f}) | {d['cpu_baseline']['value']:.3f} ({d['cpu_baseline']['kind']}, {d['cpu_baseline']['cores']} cores) | {d['clocks']['sm_mhz']:.0f} |")
open("profiles/r02_bench_table.md","w").write("# Round 2 — one bench line per BASELINE config (B200, 1 GPU, `bench.py --config cfgN`, final kernels)\n\n| config | pairs/step | pairs/s | ms/step | e2e pairs/s (host fp32 patches) | e2e pairs/s (host uint8 images) | sustained >= 3 s pairs/s | encoder GEMM TFLOP/s (frac of burst) | whole path frac of burst (algorithmic FLOPs) | attention ms/launch (share) | CPU arm pairs/s | SM MHz |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n"+"\n".join(rows)+"\n")
print(open("profiles/r02_bench_table.md").read())
PY
