#!/bin/bash
# Turns the scratch outputs of scripts/gpu_final.sh (gpurun_out/) into the tracked evidence under profiles/.
R=${1:-r01}
cd "$(dirname "$0")/.."
mkdir -p profiles
cp gpurun_out/bench.json profiles/${R}_bench.json
cp gpurun_out/bench_reference.json profiles/${R}_bench_reference.json
cp gpurun_out/launches.csv profiles/${R}_launches.csv
python scripts/launch_summary.py gpurun_out/launches.csv > profiles/${R}_launches.md
for k in gemm2 attention hbm; do
  python scripts/ncu_summary.py gpurun_out/prof_${k}.ncu-rep > profiles/${R}_ncu_${k}.txt 2>&1
done
python scripts/gemm_traffic.py gpurun_out/prof_gemm2.ncu-rep profiles/gemm_traffic.json
tail -3 gpurun_out/pytest_gpu.log > profiles/${R}_pytest_gpu.txt
tail -2 gpurun_out/smoke.log > profiles/${R}_smoke.txt
cat gpurun_out/host.txt > profiles/${R}_host.txt
python scripts/make_summary.py ${R} > /dev/null
