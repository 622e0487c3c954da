#!/bin/bash
# End-of-round evidence run: full GPU tests, smoke, bench (own + reference arm), ncu launch list + full captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "=== pytest: $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "^N=|cfg4|FAILED|^E  " gpurun_out/pytest_gpu.log | head -30
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "=== smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "=== bench rc=$?"; cat gpurun_out/bench.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
echo "=== reference arm rc=$?"; cat gpurun_out/bench_reference.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 360 -c 200 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "=== ncu launches rc=$?"
bash scripts/gpu_ncu.sh "gemm2_kernel" 60 4 prof_gemm2
bash scripts/gpu_ncu.sh "attention_kernel" 14 1 prof_attention
bash scripts/gpu_ncu.sh "layernorm_kernel|diffnet_fused|patch_gather|embed_assemble" 20 8 prof_hbm
