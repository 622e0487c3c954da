#!/bin/bash
# A/B of the CTA-pair tile width for the N = 768 projections (attn.out, fc2) at small and large M.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
out=gpurun_out/bn768_ab.jsonl; : > $out
for rep in 1 2; do
  for pairs in 32 64 256 1024; do
    for bn in 192 256; do
      GVC_SHAPES=gemm_out,gemm_fc2 GVC_VTQ_ONLY=1 VTQ_GEMM_BN_N768=$bn python scripts/gemm_vs_cublas.py $pairs >> $out 2>gpurun_out/bn768_ab.err
    done
  done
done
cat $out
