#!/bin/bash
# compute-sanitizer (memcheck) over a tiny end-to-end forward and the per-kernel tests with small shapes.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os, numpy as np, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "tests")]
import synth, vtamiq_b200
torch.manual_seed(0)
m = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False, num_keep_layers=2), cuda_graph=False).eval().cuda()
B, N, H, W = 2, 150, 96, 128
rng = np.random.default_rng(0)
imgs = torch.stack([torch.stack([synth.to_tensor_normalized(synth.make_pair(p, H, W, 0.1)[k]) for p in range(B)]) for k in range(2)]).cuda()
smp = [torch.from_numpy(np.stack([synth.jittered_samples(rng, H, W, N) for _ in range(B)])).cuda()]
with torch.no_grad():
    q = m.forward_from_images(imgs, smp)
torch.cuda.synchronize()
print("q", q.cpu().numpy())
PY
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 600 --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|q \[" gpurun_out/sanitizer_memcheck.log | head
VTQ_FUSE_LN=1 timeout 900 compute-sanitizer --tool memcheck --launch-timeout 600 --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_memcheck_fuse_ln.log 2>&1
echo "memcheck (LayerNorm folding) rc=$?"; grep -E "ERROR SUMMARY|Invalid|q \[" gpurun_out/sanitizer_memcheck_fuse_ln.log | head
