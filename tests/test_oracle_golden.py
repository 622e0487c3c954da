"""The oracle (oracle/) against the fixtures produced by the unmodified reference (tests/golden/make_golden.py).
CPU only.  This is what pins the checker that the GPU parity tests then trust."""
import ast
import os

import numpy as np
import pytest
import torch

import synth
from oracle import patch_oracle, vtamiq_oracle

PATCH_CASES = ["single", "multi3", "odd2", "clamp"]
FORWARD_CASES = ["default", "scales3", "traincfg", "adapters"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def _samples(g):
    return [g[f"samples_{i}"] for i in range(int(g["n_levels"]))]


@pytest.mark.parametrize("case", PATCH_CASES)
def test_patch_oracle_bit_exact(golden_dir, case):
    g = _load(golden_dir, "patches_" + case)
    tens = np.stack([synth.to_tensor_normalized(g["ref_u8"]).numpy(), synth.to_tensor_normalized(g["dist_u8"]).numpy()])
    patches, pos, scales = patch_oracle.extract_patches(tens, _samples(g))
    assert patches.dtype == np.float32 and pos.dtype == np.float32
    assert np.array_equal(patches.view(np.uint32), g["patches"].view(np.uint32)), "gather not bit-exact"
    assert np.array_equal(pos.view(np.uint32), g["pos"].view(np.uint32)), "uv not bit-exact"
    if "scales" in g.files:
        assert np.array_equal(scales, g["scales"])
    else:
        assert scales is None
    assert float(pos.max()) < 1.0 and float(pos.min()) >= 0.0


def test_avgpool_tree_matches_torch():
    rng = np.random.default_rng(0)
    t = rng.standard_normal((2, 3, 37, 50)).astype(np.float32)
    want = torch.nn.AvgPool2d(2)(torch.from_numpy(t)).numpy()
    got = patch_oracle.avgpool2x2(t)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_index_formulas_match_torch():
    rng = np.random.default_rng(1)
    pos = rng.random((1000, 2)).astype(np.float32)
    pos[:5] = [[0, 0], [0.99999899, 0.99999899], [1 / 24, 1 / 24], [0.5, 0.25], [23 / 24, 0.0]]
    p = torch.floor(torch.from_numpy(pos) * 24)
    want = ((p[:, 0] * 24 + p[:, 1]) + 1).to(torch.long).numpy()
    got = patch_oracle.pos_index(pos, 24)
    assert np.array_equal(got, want) and got.min() >= 1 and got.max() <= 576
    sc = np.array([0, 1, 2, 3, 7, -1], np.float32)
    assert patch_oracle.scale_index(sc, 3).tolist() == [1, 2, 3, 3, 3, 1]


def _build_mine(g):
    import vtamiq_b200
    vit_cfg = ast.literal_eval(str(g["vit_cfg"]))
    vt_kwargs = ast.literal_eval(str(g["vt_kwargs"]))
    torch.manual_seed(0)
    m = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False, **vit_cfg), **vt_kwargs).eval()
    synth.perturb_(m)
    return m


@pytest.mark.parametrize("case", FORWARD_CASES)
def test_forward_oracle_matches_reference(golden_dir, case):
    g = _load(golden_dir, "forward_" + case)
    m = _build_mine(g)
    sd = m.state_dict()
    # same seed, same construction order -> the very same weights the reference had
    assert synth.state_hash(sd) == str(g["state_hash"])
    B = int(g["B"])
    smp = [g[f"samples_{i}"] for i in range(int(g["n_levels"]))]
    P, POS, SC = [], [], []
    for b in range(B):
        tens = np.stack([synth.to_tensor_normalized(g["u8"][b, k]).numpy() for k in range(2)])
        patches, pos, scales = patch_oracle.extract_patches(tens, [s[b] for s in smp])
        P.append(patches); POS.append(pos); SC.append(scales)
    P, POS = torch.from_numpy(np.stack(P)), torch.from_numpy(np.stack(POS))
    use_sc = SC[0] is not None
    SCt = torch.from_numpy(np.stack(SC)).to(torch.float32) if use_sc else None
    q, inter = vtamiq_oracle.vtamiq_forward(
        sd, (P[:, 0], P[:, 1]), (POS[:, 0], POS[:, 1]), (SCt[:, 0], SCt[:, 1]) if use_sc else None,
        return_intermediates=True)
    for i, key in enumerate(("ref", "dist")):
        st = inter[key]
        np.testing.assert_allclose(st[0][:, [0, -1]].numpy(), g["embed_tok"][i], rtol=0, atol=1e-5)
        np.testing.assert_allclose(st[1][:, [0, -1]].numpy(), g["layer0_tok"][i], rtol=0, atol=2e-5)
    np.testing.assert_allclose(inter["diff"].numpy(), g["diff"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(q.numpy(), g["q"], rtol=0, atol=2e-5)


# ------------------------------------------------------------------------------------------ tail gradients
def _tail_oracle_grads(g, mode):
    """Oracle tail under torch.autograd on the reference's recorded d0 / DropPath factors."""
    import ast
    import vtamiq_b200
    vit_cfg, vt_kwargs = ast.literal_eval(str(g["vit_cfg"])), ast.literal_eval(str(g["vt_kwargs"]))
    torch.manual_seed(0)
    m = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False, **vit_cfg), **vt_kwargs)
    synth.perturb_(m)
    with torch.no_grad():
        gen = torch.Generator().manual_seed(9)
        for name, p in m.named_parameters():
            if name.startswith(("quality_decoder", "q_predictor")) and name.endswith("bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=gen))
    assert synth.state_hash(m.state_dict()) == str(g[f"state_hash_{mode}"])
    sd = {k: v.detach().clone().requires_grad_(k.startswith(("quality_decoder", "q_predictor", "diff_scale")))
          for k, v in m.state_dict().items()}
    cfg = vtamiq_oracle._cfg_from_state(sd)
    d0 = torch.from_numpy(g[f"d0_{mode}"])
    drop = torch.from_numpy(g[f"drop_{mode}"]) if mode == "train" else None
    q = vtamiq_oracle.diffnet_head(sd, cfg, d0 * sd["diff_scale.gamma"], drop_scale=drop)
    (q * torch.from_numpy(g["wts"])).sum().backward()
    return m, sd, q


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_oracle_tail_gradients_match_reference_autograd(golden_dir, mode):
    """The oracle's differentiable tail (incl. DropPath factors) reproduces the gradients the REFERENCE's autograd
    produced (tests/golden/make_golden.py::tail_grads_case) — this pins the checker of the CUDA backward."""
    g = np.load(os.path.join(golden_dir, "tail_grads.npz"))
    _, sd, q = _tail_oracle_grads(g, mode)
    assert np.abs(q.detach().numpy() - g[f"q_{mode}"]).max() < 1e-5
    names = [str(n) for n in g[f"names_{mode}"]]
    assert len(names) == 40 and "diff_scale.gamma" in names
    for name in names:
        want = g[f"grad_{mode}/{name}"]
        got = synth.grad_probe(sd[name].grad)
        ok = ~np.isnan(want)
        scale = max(np.abs(want[ok][2:]).max(), 1e-6)
        assert np.abs(got[ok][2:] - want[ok][2:]).max() <= 2e-5 * scale + 1e-7, name
        assert abs(got[1] - want[1]) <= 1e-4 * want[1] + 1e-6, name


# ------------------------------------------------------------------------------------------ npz loader
@pytest.mark.parametrize("tag,ntok", [("same", 577), ("zoom", 197)])
def test_load_from_matches_reference_loader(golden_dir, tag, ntok):
    """VisionTransformer.load_from on a synthetic JAX-format checkpoint gives the state_dict the REFERENCE's loader
    gave on the same file (hash over every key and tensor), including the ndimage.zoom resize of a 14x14 positional
    grid (transformer.py:428-455)."""
    import vtamiq_b200
    g = np.load(os.path.join(golden_dir, "load_from.npz"))
    torch.manual_seed(0)
    m = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False)).eval()
    m.transformer.load_from(synth.synthetic_vit_npz(seed=21, pos_tokens=ntok), True, True)
    sd = m.state_dict()
    pos = sd["transformer.embeddings.positional_embeddings.positional_embeddings"].numpy()[0, ::7]
    assert np.array_equal(pos, g[f"pos_{tag}"])
    assert synth.state_hash(sd) == str(g[f"hash_{tag}"])
