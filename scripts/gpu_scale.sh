#!/bin/bash
# Multi-GPU lines on ONE box with N GPUs: strong scaling (cfg5, one global batch of 2048 pairs split over the ranks) and
# the default weak-scaling line (cfg2, 32 pairs per GPU).  usage: gpu_scale.sh N
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
run () {  # $1 = output name, rest = bench args
  local out=$1; shift
  if [ "$N" = 1 ]; then timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/$out 2> gpurun_out/$out.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
         bench.py --gpus $N "$@" > gpurun_out/$out 2> gpurun_out/$out.err; fi
  echo "=== $out rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$out").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["scaling"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "per-rank ms", d.get("per_rank_ms"), "straggler", d.get("straggler_rank"), "e2e", d["e2e"]["value"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/$out.err").read()[-800:])
PY
}
run scale_strong_cfg5_${N}gpu.json --config cfg5 --global-pairs 2048 --steps 8 --warmup 3 --no-cpu --no-sustained
run scale_weak_cfg2_${N}gpu.json --steps 30 --warmup 5 --no-cpu --no-sustained
