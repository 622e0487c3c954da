#!/bin/bash
# Round 2, run M: new get_iqa_patches branch tests; ncu --set full of the gather / pyramid kernels (cfg3, uint8 and fp32
# images), of attention at cfg4 (S = 5001) and of the encoder GEMMs at cfg2.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "get_iqa_patches or patch_gather" 2>&1 | tail -3
ncu_cap () {  # $1 regex $2 skip $3 count $4 name, rest: bench args
  local re=$1 sk=$2 ct=$3 nm=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$re" -s "$sk" -c "$ct" -o "gpurun_out/$nm" -f \
     python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-sustained "$@" > "gpurun_out/ncu_$nm.log" 2>&1
  echo "=== ncu $nm rc=$?"
}
ncu_cap "patch_gather|avgpool2x2|normalize_u8" 0 12 prof_gather_cfg3_u8 --config cfg3
ncu_cap "patch_gather|avgpool2x2|normalize_u8" 0 12 prof_gather_cfg3_f32 --config cfg3 --images fp32
ncu_cap "attention_kernel" 3 1 prof_attention_cfg4 --config cfg4
ncu_cap "gemm2_kernel" 60 4 prof_gemm2 --config cfg2
ls -la gpurun_out/*.ncu-rep
