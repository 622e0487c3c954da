"""On-box GPU comparison bar: the UNMODIFIED reference model (oracle/_ref) run by stock PyTorch eager on
the same B200 — fp16 autocast like the reference's own GPU path (train.py:514,:602: cuBLASLt GEMMs, unfused
matmul / softmax / matmul attention, separate ref and dist passes) and plain fp32 — next to vtamiq_b200 on the same
patch tensors already resident in HBM.  Context only: the tier's reference arm is the CPU forward (bench.py --impl
reference); nothing of this is on the product path.

    python scripts/eager_gpu_reference.py [pairs] [patches]     -> one JSON line
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

import synth  # noqa: E402
import vtamiq_b200  # noqa: E402
from oracle import reference_runner  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500
dev = torch.device("cuda:0")
ref = reference_runner.build_model(perturb=synth.perturb_).to(dev)
torch.manual_seed(0)
ours = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False)).eval()
synth.perturb_(ours)
ours = ours.to(dev)
g = torch.Generator(device=dev).manual_seed(1)
patches = [torch.randn(B, N, 3, 16, 16, device=dev, generator=g) for _ in range(2)]
pos = [torch.rand(B, N, 2, device=dev, generator=g) * 0.999 for _ in range(2)]


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        q = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, q


def ref_fp16():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        return ref((patches[0], patches[1]), (pos[0], pos[1]), (None, None))[0]


def ref_fp32():
    with torch.no_grad():
        return ref((patches[0], patches[1]), (pos[0], pos[1]), (None, None))[0]


def mine():
    with torch.no_grad():
        return ours((patches[0], patches[1]), (pos[0], pos[1]), None)[0]


torch.backends.cuda.matmul.allow_tf32 = False
ms32, q32 = timed(ref_fp32, 5)
ms16, q16 = timed(ref_fp16, 10)
msb, qb = timed(mine, 20)
print(json.dumps({
    "what": "stock PyTorch eager forward of the unmodified reference on the same B200 vs vtamiq_b200 (patch tensors "
            "resident in HBM, through VTAMIQ.forward)",
    "pairs": B, "patches": N,
    "reference_eager_fp32": {"ms": round(ms32, 3), "pairs_per_s": round(B / ms32 * 1e3, 1)},
    "reference_eager_fp16_autocast": {"ms": round(ms16, 3), "pairs_per_s": round(B / ms16 * 1e3, 1),
                                      "max_abs_dq_vs_fp32": float((q16.float() - q32).abs().max())},
    "vtamiq_b200_fp16": {"ms": round(msb, 3), "pairs_per_s": round(B / msb * 1e3, 1),
                         "max_abs_dq_vs_fp32": float((qb - q32.reshape(qb.shape)).abs().max())},
    "speedup_vs_eager_fp16": round(ms16 / msb, 2),
}))
