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
# round 2: the device sampler, the training tail (forward + backward) and the unaligned gather under memcheck
cat > /tmp/san2.py <<'PY'
import sys, os, numpy as np, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "tests")]
import synth, vtamiq_b200
from vtamiq_b200.patch_sampling import sample_batch, extract_patches
torch.manual_seed(0)
g = torch.Generator(device="cuda").manual_seed(5)
lv = sample_batch(3, 200, 264, 90, 16, 2, 2.0, device="cuda", generator=g)
print("sampler", [tuple(t.shape) for t in lv], float(lv[0].max()))
tens = torch.randn(2, 3, 200, 264, device="cuda")
smp = [np.stack([lv[s][k].cpu().numpy() for k in range(2)]) for s in range(len(lv))]   # one set per image
p, pos, sc = extract_patches(tens, smp)
print("unaligned gather", tuple(p.shape), float(pos.max()))
m = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False, num_keep_layers=1), num_rgs=1, num_rcabs=2).cuda()
m.set_freeze_state(True, dict(freeze_dict_vit=dict(freeze_encoder=True, freeze_encoder_adapters=True, freeze_encoder_layerscale=True,
    freeze_embeddings_patch=True, freeze_embeddings_cls_token=True, freeze_embeddings_extra_tokens=True, freeze_embeddings_pos=True,
    freeze_embeddings_scale=True), freeze_quality_decoder=False, freeze_q_predictor=False, freeze_w_predictor=False)) if hasattr(m, "set_freeze_state") else None
for prm in m.transformer.parameters(): prm.requires_grad = False
m.train()
B, N = 3, 40
patches = [torch.randn(B, N, 3, 16, 16, device="cuda") for _ in range(2)]
pos = [torch.rand(B, N, 2, device="cuda") * 0.99 for _ in range(2)]
q, _ = m((patches[0], patches[1]), (pos[0], pos[1]), (None, None))
q.sum().backward()
torch.cuda.synchronize()
print("tail backward", float(q.sum()), sum(float(p.grad.abs().sum()) for p in m.parameters() if p.grad is not None) > 0)
PY
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 600 --print-limit 20 python /tmp/san2.py > gpurun_out/sanitizer_memcheck_r2.log 2>&1
echo "memcheck (round-2 paths) rc=$?"; grep -E "ERROR SUMMARY|Invalid|sampler|unaligned|tail backward|Error" gpurun_out/sanitizer_memcheck_r2.log | head
