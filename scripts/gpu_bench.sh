#!/bin/bash
# Round driver for one gpurun call: full GPU test suite, bench line, ncu launch list, ncu full capture of the GEMM.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
if [ "$1" != "notest" ]; then
  timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
  echo "=== pytest: $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "^N=|FAILED|^E  " gpurun_out/pytest_gpu.log | head -30
fi
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "=== bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" != "noncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 576 -c 300 --csv \
     --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu > gpurun_out/ncu_launch.log 2>&1
  echo "=== ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 60 -c 4 \
     -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 1 --no-graph --no-cpu > gpurun_out/ncu_gemm.log 2>&1
  echo "=== ncu gemm rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 12 -c 1 \
     -o gpurun_out/prof_attn -f python bench.py --steps 1 --warmup 1 --no-graph --no-cpu > gpurun_out/ncu_attn.log 2>&1
  echo "=== ncu attn rc=$?"
fi
