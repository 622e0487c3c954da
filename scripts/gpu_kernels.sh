#!/bin/bash
# Runs the per-kernel GPU tests in separate processes (a trapped kernel poisons its CUDA context).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/smi.txt 2>&1
run() {  # name, -k expression, file
  timeout 600 python -m pytest "$3" -q -m gpu -k "$2" -p no:cacheprovider 2>&1 | tail -60 > "gpurun_out/$1.log"
  echo "=== $1: $(tail -1 gpurun_out/$1.log)"
}
run k_gemm "gemm" tests/test_gpu_kernels.py
run k_attn "attention" tests/test_gpu_kernels.py
run k_rows "layernorm or cast or embed or cls_diff or diffnet" tests/test_gpu_kernels.py
run k_patch "patch or avgpool" tests/test_gpu_kernels.py
if [ "$1" = "all" ]; then
  run fwd "test_" tests/test_gpu_forward.py
fi
for f in k_gemm k_attn k_rows k_patch fwd; do [ -f gpurun_out/$f.log ] && { echo "----- $f"; grep -E "^(FAILED|ERROR|E  )|passed|failed|error" gpurun_out/$f.log | head -40; }; done
