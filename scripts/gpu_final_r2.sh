#!/bin/bash
# End-of-round-2 evidence run: full GPU tests, smoke, bench lines for every BASELINE config (own arm) + the reference arm,
# ncu launch list of one un-graphed cfg2 step.  Outputs under gpurun_out/final_r2/ (scripts/make_profiles_r2.sh copies
# the tracked subset into profiles/).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; O=gpurun_out/final_r2; mkdir -p $O
nproc > $O/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $O/host.txt; nvidia-smi -L >> $O/host.txt
# DRAM traffic of the four encoder projections from a fresh --set full capture (bench.py reports it as roofline.traffic)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm2_kernel" -s 60 -c 4 -o gpurun_out/prof_gemm2 -f \
   python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-sustained > $O/ncu_gemm2.log 2>&1; echo "=== ncu gemm2 rc=$?"
python scripts/gemm_traffic.py gpurun_out/prof_gemm2.ncu-rep profiles/gemm_traffic.json && cp profiles/gemm_traffic.json $O/gemm_traffic.json
python scripts/ncu_summary.py gpurun_out/prof_gemm2.ncu-rep > $O/ncu_gemm2.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider -s --durations=10 2>&1 | tail -60 > $O/pytest_gpu.log
echo "=== pytest: $(grep -E 'passed|failed|error' $O/pytest_gpu.log | tail -1)"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "=== smoke rc=$?"; tail -1 $O/smoke.log | cut -c1-300
timeout 900 python bench.py --steps 30 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "=== bench cfg2 rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "=== reference arm rc=$?"; cut -c1-400 $O/bench_reference.json
for cfg in cfg1 cfg3 cfg4 cfg5; do
  timeout 900 python bench.py --config $cfg --steps 20 --warmup 5 > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "=== bench $cfg rc=$?"
done
for cfg in cfg2 cfg1 cfg3 cfg4 cfg5; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$cfg.json").read()); r=d["roofline"]; s=d["sustained"]
    print("$cfg", d["value"], "pairs/s", d["ms_per_step"], "ms/step e2e", d["e2e"]["value"], "e2e_img", d["e2e_from_images"]["value"], "gemm", r["achieved"], r["frac_of_burst"], "algo frac burst", d["frac_of_bf16_peak"]["burst"], "sustained", s and s["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$cfg parse failed", e)
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 360 -c 200 --csv \
   --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu --no-sustained > $O/ncu_launch.log 2>&1
echo "=== ncu launches rc=$?"
