"""Generates the committed golden fixtures by running the UNMODIFIED reference (ch-andrei/VTAMIQ at
/root/reference) on CPU fp32.  Only runs in the build container (the reference does not travel); the
fixtures it writes are what pins oracle/ on every other machine.

    python tests/golden/make_golden.py

Needs oracle/ref_shims on the path for `timm`, `matplotlib`, `skimage` (absent from the image; none of them
does inference arithmetic).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VTAMIQ_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_shims"), REF, os.path.join(ROOT, "tests"), ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import synth  # noqa: E402
from data.patch_sampling import GRID_TYPE_PERTURBED_SIMPLE, PatchSampler, get_iqa_patches  # noqa: E402
from modules.vtamiq.vtamiq import VTAMIQ  # noqa: E402

torch.set_num_threads(8)


def sampler():
    # train_config.py:284-288
    return PatchSampler(centerbias_weight=0.0, diff_weight=0.0, uniform_weight=0.1, grid_type=GRID_TYPE_PERTURBED_SIMPLE)


def capture_samples(fn):
    """Run fn while recording what stratified_grid_sampling returned (the coordinates are internal to
    get_iqa_patches; the device gather needs them as inputs)."""
    import data.patch_sampling as ps
    rec = []
    orig = ps.stratified_grid_sampling

    def spy(*a, **k):
        out = orig(*a, **k)
        rec.append(np.array(out, dtype=np.float64, copy=True))
        return out
    ps.stratified_grid_sampling = spy
    try:
        res = fn()
    finally:
        ps.stratified_grid_sampling = orig
    return res, rec


def patches_case(name, H, W, N, n_scales, ratio, seed):
    ref, dist = synth.make_pair(seed, H, W, 0.1)
    tens = (synth.to_tensor_normalized(ref), synth.to_tensor_normalized(dist))
    (patches, pos, scales), rec = capture_samples(lambda: get_iqa_patches(
        (ref, dist), tens, N, 16, sampler(), n_scales, scale_num_samples_ratio=ratio,
        use_aligned_patches=True, random_seed=seed))
    out = dict(ref_u8=ref, dist_u8=dist, N=N, n_scales_requested=n_scales, ratio=ratio,
               patches=patches.numpy(), pos=pos.numpy())
    if scales is not None:
        out["scales"] = scales.numpy()
    for i, s in enumerate(rec):
        out[f"samples_{i}"] = s
    out["n_levels"] = len(rec)
    np.savez_compressed(os.path.join(HERE, f"patches_{name}.npz"), **out)
    print("patches", name, patches.shape, [s.shape for s in rec], None if scales is None else np.bincount(scales[0].numpy()))


def patches_branch_case(name, H, W, N, n_scales, ratio, seed, aligned, shuffle, sampler_kw=None):
    """The non-default branches of get_iqa_patches (patch_sampling.py:506-508 slot permutation, :530-531/:561 one
    coordinate set per image, :46-222/:603-605 difference- / centre-bias-weighted sampling with the weight map pooled
    per level).  Recorded: what the sampler returned for every draw, the weight map it was handed at every draw and
    compute_diff's output, so the device path can be replayed through vtamiq_b200.get_iqa_patches with a stub
    sampler."""
    import data.patch_sampling as ps
    ref, dist = synth.make_pair(seed, H, W, 0.1)
    tens = (synth.to_tensor_normalized(ref), synth.to_tensor_normalized(dist))
    smp = PatchSampler(**sampler_kw) if sampler_kw else sampler()
    draws, diffs = [], []
    orig = smp.get_sample_params

    def spy(h, w, ho, wo, diff=None, num_samples=1, debug=False):
        out = orig(h, w, ho, wo, diff=diff, num_samples=num_samples, debug=debug)
        draws.append(np.array(out, dtype=np.float64, copy=True))
        diffs.append(None if diff is None else np.array(diff, copy=True))
        return out
    smp.get_sample_params = spy
    from PIL import Image
    imgs = (Image.fromarray(ref), Image.fromarray(dist))
    diff0 = smp.compute_diff(imgs)
    patches, pos, scales = get_iqa_patches(imgs, tens, N, 16, smp, n_scales, scale_num_samples_ratio=ratio,
                                           use_aligned_patches=aligned, randomize_patch_scale_order=shuffle,
                                           random_seed=seed)
    out = dict(ref_u8=ref, dist_u8=dist, N=N, n_scales_requested=n_scales, ratio=ratio, seed=seed, aligned=aligned,
               shuffle=shuffle, patches=patches.numpy(), pos=pos.numpy(), n_draws=len(draws),
               has_diff=diff0 is not None)
    if scales is not None:
        out["scales"] = scales.numpy()
    if diff0 is not None:
        out["diff0"] = np.asarray(diff0)
    for i, (d, w) in enumerate(zip(draws, diffs)):
        out[f"draw_{i}"] = d
        if w is not None:
            out[f"weight_{i}"] = w
    np.savez_compressed(os.path.join(HERE, f"patches_{name}.npz"), **out)
    print("patches", name, patches.shape, len(draws), "draws", None if scales is None else np.bincount(scales[0].numpy()))


def forward_case(name, vit_cfg, vt_kwargs, B, H, W, N, n_scales, ratio):
    torch.manual_seed(0)
    model = VTAMIQ(vit_config=dict(pretrained=False, **vit_cfg), **vt_kwargs).eval()
    synth.perturb_(model)
    levels = synth.graded_levels(B)
    P, POS, SC, U8, SMP = [], [], [], [], []
    for p in range(B):
        ref, dist = synth.make_pair(p, H, W, float(levels[p]))
        tens = (synth.to_tensor_normalized(ref), synth.to_tensor_normalized(dist))
        (patches, pos, scales), rec = capture_samples(lambda: get_iqa_patches(
            (ref, dist), tens, N, 16, sampler(), n_scales, scale_num_samples_ratio=ratio,
            use_aligned_patches=True, random_seed=p))
        P.append(patches); POS.append(pos); SC.append(scales); U8.append(np.stack([ref, dist])); SMP.append(rec)
    patches = torch.stack(P)            # (B, 2, N, 3, 16, 16)
    pos = torch.stack(POS)
    use_sc = SC[0] is not None
    scales = torch.stack(SC).to(torch.float32) if use_sc else None   # train.py:254 casts everything to fp32
    hooks = {}
    vit = model.transformer

    def keep(tag):
        def fn(mod, inp, out):
            hooks.setdefault(tag, []).append((out[0] if isinstance(out, tuple) else out).detach().clone())
        return fn
    # Encoder / EncoderLayer are invoked through .forward() (no hooks fire): tap the LayerNorms instead —
    # the input of layers[1].attention_norm is the residual stream after block 0.
    def keep_in(tag):
        def fn(mod, inp):
            hooks.setdefault(tag, []).append(inp[0].detach().clone())
        return fn
    h1 = vit.embeddings.register_forward_hook(keep("embed"))
    h2 = vit.encoder.layers[1].attention_norm.register_forward_pre_hook(keep_in("layer0"))
    h3 = vit.encoder.encoder_norm.register_forward_hook(keep("encoded"))
    h4 = model.diff_scale.register_forward_hook(keep("diff"))
    with torch.no_grad():
        q, _ = model((patches[:, 0].clone(), patches[:, 1].clone()), (pos[:, 0].clone(), pos[:, 1].clone()),
                     (scales[:, 0].clone(), scales[:, 1].clone()) if use_sc else (None, None))
    for h in (h1, h2, h3, h4):
        h.remove()
    out = dict(
        q=q.numpy(), state_hash=synth.state_hash(model.state_dict()), u8=np.stack(U8), levels=levels,
        B=B, N=N, n_scales_requested=n_scales, ratio=ratio,
        vit_cfg=repr(vit_cfg), vt_kwargs=repr(vt_kwargs),
        # sparse probes of intermediates, ref stream then dist stream: token 0 and the last token
        embed_tok=np.stack([hooks["embed"][i][:, [0, -1]].numpy() for i in range(2)]),
        layer0_tok=np.stack([hooks["layer0"][i][:, [0, -1]].numpy() for i in range(2)]),
        encoded_cls=np.stack([hooks["encoded"][i][:, 0].numpy() for i in range(2)]),
        diff=hooks["diff"][0].numpy(),
        n_levels=len(SMP[0]),
    )
    for lvl in range(len(SMP[0])):
        out[f"samples_{lvl}"] = np.stack([SMP[p][lvl] for p in range(B)])   # (B, 2, n_lvl)
    np.savez_compressed(os.path.join(HERE, f"forward_{name}.npz"), **out)
    print("forward", name, "q =", q.numpy())


def correlations_case():
    """utils/misc/correlations.py:21-51 on seeded score vectors (one with ties, one already in [0,1])."""
    from utils.misc.correlations import compute_correlations
    rng = np.random.default_rng(11)
    out = {}
    for k, (n, ties) in enumerate([(60, False), (400, True), (25, False)]):
        a = rng.normal(size=n) * 1.3 + 4.0
        b = 0.8 * a + rng.normal(size=n) * (0.25 + 0.2 * k) + 0.1 * a ** 2
        if ties:
            a, b = np.round(a, 1), np.round(b, 1)
        c = compute_correlations(a, b)
        out[f"a{k}"], out[f"b{k}"] = a, b
        out[f"want{k}"] = np.array([c[key] for key in ("SROCC", "KROCC", "PLCC", "RMSE", "PLCC_NOFIT", "RMSE_NOFIT")])
        print("correlations", k, out[f"want{k}"])
    np.savez_compressed(os.path.join(HERE, "correlations.npz"), **out)


def sampler_draws_case():
    """Raw draws of the reference's default coordinate sampler (PERTURBED_SIMPLE): the statistical yardstick for
    vtamiq_b200.patch_sampling.perturbed_grid_samples."""
    out = {}
    for name, (h, w, n, draws) in dict(cfg2=(384, 512, 500, 24), small=(96, 128, 64, 100), tall=(300, 100, 37, 100)).items():
        np.random.seed(1234)
        smp = sampler()
        out[name] = np.stack([smp.get_sample_params(h, w, 16, 16, num_samples=n) for _ in range(draws)]).astype(np.float32)
        out[name + "_hwn"] = np.array([h, w, n])
        print("sampler draws", name, out[name].shape)
    np.savez_compressed(os.path.join(HERE, "sampler_draws.npz"), **out)


def load_from_case():
    """The reference's own npz loader (transformer.py:643-668 -> :287-325, :428-455) on a synthetic JAX-format
    checkpoint: once with a 577-token positional table (copied as is) and once with a 197-token one (14x14 grid ->
    ndimage.zoom to 24x24).  The fixture keeps the state_dict hash and the complete resized positional table."""
    out = {}
    for tag, ntok in (("same", 577), ("zoom", 197)):
        torch.manual_seed(0)
        model = VTAMIQ(vit_config=dict(pretrained=False)).eval()
        w = synth.synthetic_vit_npz(seed=21, pos_tokens=ntok)
        model.transformer.load_from(w, True, True)
        sd = model.state_dict()
        out[f"hash_{tag}"] = synth.state_hash(sd)
        out[f"pos_{tag}"] = sd["transformer.embeddings.positional_embeddings.positional_embeddings"].numpy()[0, ::7].copy()
        print("load_from", tag, out[f"hash_{tag}"][:16])
    np.savez_compressed(os.path.join(HERE, "load_from.npz"), **out)


def tail_grads_case():
    """Gradients the REFERENCE's autograd gives the parameters behind the encoder (diff_scale, quality_decoder,
    q_predictor) for loss = sum_b w_b q_b: evaluation mode (DropPath identity) and training mode (DropPath masks
    drawn from torch's CPU generator after manual_seed(77): the fixture records them through a hook).
    The encoder is frozen like set_freeze_state does (backbone.py:62-106); its output difference d0 is recorded so
    the tail can be re-run on its own."""
    vit_cfg, vt_kwargs = dict(num_keep_layers=1), dict(num_rgs=2, num_rcabs=2)
    B, N = 5, 12
    g = torch.Generator().manual_seed(3)
    patches = [torch.randn(B, N, 3, 16, 16, generator=g) for _ in range(2)]
    pos = [torch.rand(B, N, 2, generator=g) * 0.999 for _ in range(2)]
    wts = torch.linspace(0.5, 1.5, B)
    out = dict(vit_cfg=repr(vit_cfg), vt_kwargs=repr(vt_kwargs), B=B, wts=wts.numpy())
    for mode in ("eval", "train"):
        torch.manual_seed(0)
        model = VTAMIQ(vit_config=dict(pretrained=False, **vit_cfg), **vt_kwargs)
        synth.perturb_(model)
        with torch.no_grad():   # make every tail parameter class matter (biases, PReLU slopes)
            gen = torch.Generator().manual_seed(9)
            for name, p in model.named_parameters():
                if name.startswith(("quality_decoder", "q_predictor")) and name.endswith("bias"):
                    p.copy_(0.05 * torch.randn(p.shape, generator=gen))
        for p in model.transformer.parameters():
            p.requires_grad = False
        model.train(mode == "train")
        rec, masks = {}, []
        h = model.diff_scale.register_forward_pre_hook(lambda m, inp: rec.__setitem__("d0", inp[0].detach().clone()))
        hooks = [h]
        for grp in list(model.quality_decoder)[:-1]:
            def post(mod, inp, outp, grp=grp):
                # DropPath output / input = the per-sample factor (mask / keep_prob)
                ratio = (outp / inp[0]).detach()
                masks.append(torch.nan_to_num(ratio[:, 0, 0], nan=0.0))
            hooks.append(grp.drop.register_forward_hook(post))
        torch.manual_seed(77)
        q, _ = model((patches[0], patches[1]), (pos[0], pos[1]), (None, None))
        (q * wts).sum().backward()
        for hk in hooks:
            hk.remove()
        out[f"state_hash_{mode}"] = synth.state_hash(model.state_dict())
        out[f"d0_{mode}"] = rec["d0"].numpy()
        out[f"q_{mode}"] = q.detach().numpy()
        out[f"drop_{mode}"] = torch.stack(masks).numpy()
        names = []
        for name, p in model.named_parameters():
            if p.grad is not None:
                names.append(name)
                out[f"grad_{mode}/{name}"] = synth.grad_probe(p.grad)
        out[f"names_{mode}"] = np.array(names)
        print("tail grads", mode, len(names), "params; drop factors", out[f"drop_{mode}"].tolist())
    np.savez_compressed(os.path.join(HERE, "tail_grads.npz"), **out)


if __name__ == "__main__":
    if "--load-from-only" in sys.argv:
        load_from_case()
        sys.exit(0)
    if "--tail-grads-only" in sys.argv:
        tail_grads_case()
        sys.exit(0)
    if "--correlations-only" in sys.argv:
        correlations_case()
        sys.exit(0)
    if "--adapters-only" in sys.argv:
        forward_case("adapters", dict(num_keep_layers=3, num_adapters=1, use_layer_scale=True), dict(num_rgs=1, num_rcabs=1),
                     B=3, H=96, W=128, N=64, n_scales=1, ratio=2.0)
        sys.exit(0)
    if "--branches-only" in sys.argv:
        patches_branch_case("unaligned", 128, 160, 48, 2, 2.0, seed=7, aligned=False, shuffle=False)
        # randomize_patch_scale_order=True cannot be pinned: under this image's torch (2.11) the reference itself raises
        # at patch_sampling.py:534 (index_put of float64 positions into a float32 tensor)
        patches_branch_case("weighted", 128, 192, 40, 2, 2.0, seed=9, aligned=True, shuffle=False,
                            sampler_kw=dict(centerbias_weight=0.0, diff_weight=0.6, uniform_weight=0.1, grid_type=1))  # GRID_TYPE_PERTURBED
        sys.exit(0)
    if "--sampler-only" in sys.argv:
        sampler_draws_case()
        sys.exit(0)
    patches_case("single", 96, 128, 64, 1, 2.0, seed=3)
    patches_case("multi3", 256, 256, 100, 3, 2.0, seed=4)
    patches_case("odd2", 250, 301, 80, 2, 1.75, seed=5)
    patches_case("clamp", 72, 200, 40, 3, 2.0, seed=6)      # image too small for 3 levels: clamped (:398-411)
    forward_case("default", {}, {}, B=4, H=96, W=128, N=64, n_scales=1, ratio=2.0)
    forward_case("scales3", dict(num_scales=3), {}, B=2, H=256, W=256, N=100, n_scales=3, ratio=2.0)
    forward_case("traincfg", dict(num_keep_layers=6, num_extra_tokens=8, use_layer_scale=True),
                 dict(ca_reduction=16), B=2, H=96, W=128, N=64, n_scales=1, ratio=2.0)
    forward_case("adapters", dict(num_keep_layers=3, num_adapters=1, use_layer_scale=True), dict(num_rgs=1, num_rcabs=1),
                 B=3, H=96, W=128, N=64, n_scales=1, ratio=2.0)
    patches_branch_case("unaligned", 128, 160, 48, 2, 2.0, seed=7, aligned=False, shuffle=False)
    patches_branch_case("weighted", 128, 192, 40, 2, 2.0, seed=9, aligned=True, shuffle=False,
                        sampler_kw=dict(centerbias_weight=0.0, diff_weight=0.6, uniform_weight=0.1, grid_type=1))  # GRID_TYPE_PERTURBED
    correlations_case()
    sampler_draws_case()
    load_from_case()
    tail_grads_case()
