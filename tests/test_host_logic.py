"""Host-side logic: per-scale patch budgets, scale clamp, module contract (state_dict, npz loader), sharding."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import vtamiq_b200
from vtamiq_b200 import patch_sampling as ps
from vtamiq_b200.parallel import shard_pairs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_patches_per_scale_known_values():
    # probes of the reference's compute_num_patches_per_scale
    assert ps.compute_num_patches_per_scale(500, 3, 2.0)[::-1].tolist() == [380, 96, 24]
    assert ps.compute_num_patches_per_scale(500, 3, 1.75)[::-1].tolist() == [360, 108, 32]
    assert ps.compute_num_patches_per_scale(100, 3, 2.0)[::-1].tolist() == [75, 20, 5]   # golden multi3
    assert ps.compute_num_patches_per_scale(80, 2, 1.75)[::-1].tolist() == [61, 19]      # golden odd2
    assert ps.compute_num_patches_per_scale(256, 1, 2.0).tolist() == [256]
    for n in (3, 7, 500, 5000):
        for k in (1, 2, 3, 4):
            if n >= k:
                assert ps.compute_num_patches_per_scale(n, k, 1.7).sum() == n


def test_scale_clamp_by_image_size():
    assert ps.compute_patch_num_scales(1, 384, 512, 16, 16) == 1
    assert ps.compute_patch_num_scales(3, 1024, 1024, 16, 16) == 3
    assert ps.compute_patch_num_scales(3, 256, 256, 16, 16) == 3
    assert ps.compute_patch_num_scales(3, 72, 200, 16, 16) == 2      # golden "clamp" case: 2 levels
    assert ps.compute_patch_num_scales(5, 40, 40, 16, 16) == 1


def test_shard_pairs_partition():
    for n in (0, 1, 7, 32, 2048):
        for w in (1, 2, 3, 8):
            spans = [shard_pairs(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_pairs(4, 2, 2)


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    return vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False)).eval()


def test_state_dict_layout(model):
    sd = model.state_dict()
    assert len(sd) == 326 and sum(v.numel() for v in sd.values()) == 101014674   # SURVEY §8b probe
    shapes = {
        "transformer.embeddings.cls_token": (1, 1, 768),
        "transformer.embeddings.patch_embeddings.weight": (768, 3, 16, 16),
        "transformer.embeddings.positional_embeddings.positional_embeddings": (1, 577, 768),
        "transformer.encoder.encoder_norm.weight": (768,),
        "transformer.encoder.layers.11.attn.query.weight": (768, 768),
        "transformer.encoder.layers.0.ffn.fc1.weight": (3072, 768),
        "transformer.encoder.layers.0.ffn.fc2.weight": (768, 3072),
        "diff_scale.gamma": (768,),
        "quality_decoder.0.body.0.body.1.weight": (1,),
        "quality_decoder.3.body.3.body.2.weight": (768, 768, 1),
        "quality_decoder.0.body.0.body.4.conv_du.1.weight": (96, 768, 1),
        "quality_decoder.0.body.0.body.4.conv_du.4.weight": (768, 96, 1),
        "quality_decoder.2.body.4.weight": (768, 768, 1),
        "quality_decoder.4.bias": (768,),
        "q_predictor.1.weight": (192, 768),
        "q_predictor.2.weight": (1,),
        "q_predictor.4.weight": (1, 192),
    }
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == s, k
    assert len(model.transformer.encoder.layers) == 12 and model.vit_num_layers == 12
    assert "_engine" not in "".join(sd.keys())


def test_traincfg_variant_keys():
    torch.manual_seed(0)
    m = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False, num_keep_layers=6, num_extra_tokens=8,
                                           use_layer_scale=True, num_scales=3), ca_reduction=16)
    sd = m.state_dict()
    assert tuple(sd["transformer.embeddings.extra_tokens"].shape) == (1, 8, 768)
    assert tuple(sd["transformer.embeddings.scale_embeddings.scale_embeddings"].shape) == (1, 4, 768)
    assert tuple(sd["transformer.encoder.layers.5.ls2.gamma"].shape) == (768,)
    assert "transformer.encoder.layers.6.ls1.gamma" not in sd
    assert tuple(sd["quality_decoder.1.body.2.body.4.conv_du.1.weight"].shape) == (48, 768, 1)
    m2 = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False, num_keep_layers=6, num_extra_tokens=8,
                                            use_layer_scale=True, num_scales=3), ca_reduction=16)
    m2.load_state_dict(sd, strict=True)


def test_set_freeze_state(model):
    fd = dict(freeze_dict_vit=dict(freeze_encoder=True, freeze_encoder_adapters=False, freeze_encoder_layerscale=False,
                                   freeze_embeddings_cls_token=True, freeze_embeddings_extra_tokens=True,
                                   freeze_embeddings_patch=True, freeze_embeddings_pos=True,
                                   freeze_embeddings_scale=True),
              freeze_quality_decoder=False, freeze_q_predictor=False)
    model.set_freeze_state(True, fd)
    assert not model.transformer.encoder.layers[0].attn.query.weight.requires_grad
    assert not model.transformer.embeddings.cls_token.requires_grad
    assert model.q_predictor[1].weight.requires_grad
    model.set_freeze_state(False, fd)
    assert model.transformer.encoder.layers[0].attn.query.weight.requires_grad


def test_npz_loader_layout(model, tmp_path):
    """Synthetic JAX-format checkpoint -> torch layout rules of transformer.py:287-325,:643-668."""
    rng = np.random.default_rng(0)
    H, M = 768, 3072
    w = {"embedding/kernel": rng.standard_normal((16, 16, 3, H), dtype=np.float32),
         "embedding/bias": rng.standard_normal(H, dtype=np.float32),
         "cls": rng.standard_normal((1, 1, H), dtype=np.float32),
         "Transformer/posembed_input/pos_embedding": rng.standard_normal((1, 577, H), dtype=np.float32),
         "Transformer/encoder_norm/scale": rng.standard_normal(H, dtype=np.float32),
         "Transformer/encoder_norm/bias": rng.standard_normal(H, dtype=np.float32)}
    for i in range(12):
        r = f"Transformer/encoderblock_{i}/"
        for n in ("query", "key", "value"):
            w[r + f"MultiHeadDotProductAttention_1/{n}/kernel"] = rng.standard_normal((H, 12, 64), dtype=np.float32)
            w[r + f"MultiHeadDotProductAttention_1/{n}/bias"] = rng.standard_normal((12, 64), dtype=np.float32)
        w[r + "MultiHeadDotProductAttention_1/out/kernel"] = rng.standard_normal((12, 64, H), dtype=np.float32)
        w[r + "MultiHeadDotProductAttention_1/out/bias"] = rng.standard_normal(H, dtype=np.float32)
        w[r + "MlpBlock_3/Dense_0/kernel"] = rng.standard_normal((H, M), dtype=np.float32)
        w[r + "MlpBlock_3/Dense_0/bias"] = rng.standard_normal(M, dtype=np.float32)
        w[r + "MlpBlock_3/Dense_1/kernel"] = rng.standard_normal((M, H), dtype=np.float32)
        w[r + "MlpBlock_3/Dense_1/bias"] = rng.standard_normal(H, dtype=np.float32)
        for n in ("LayerNorm_0", "LayerNorm_2"):
            w[r + n + "/scale"] = rng.standard_normal(H, dtype=np.float32)
            w[r + n + "/bias"] = rng.standard_normal(H, dtype=np.float32)
    path = tmp_path / "vit.npz"
    np.savez(path, **w)
    model.transformer.load_from(np.load(path))
    sd = model.state_dict()
    L = "transformer.encoder.layers.7."
    r = "Transformer/encoderblock_7/"
    eq = lambda a, b: np.array_equal(a.numpy(), b)
    assert eq(sd[L + "attn.query.weight"], w[r + "MultiHeadDotProductAttention_1/query/kernel"].reshape(H, H).T)
    assert eq(sd[L + "attn.out.weight"], w[r + "MultiHeadDotProductAttention_1/out/kernel"].reshape(H, H).T)
    assert eq(sd[L + "attn.key.bias"], w[r + "MultiHeadDotProductAttention_1/key/bias"].reshape(-1))
    assert eq(sd[L + "ffn.fc1.weight"], w[r + "MlpBlock_3/Dense_0/kernel"].T)
    assert eq(sd[L + "ffn.fc2.weight"], w[r + "MlpBlock_3/Dense_1/kernel"].T)
    assert eq(sd[L + "ffn_norm.weight"], w[r + "LayerNorm_2/scale"])
    assert eq(sd["transformer.embeddings.patch_embeddings.weight"], w["embedding/kernel"].transpose(3, 2, 0, 1))
    assert eq(sd["transformer.embeddings.positional_embeddings.positional_embeddings"],
              w["Transformer/posembed_input/pos_embedding"])
    assert eq(sd["transformer.embeddings.cls_token"], w["cls"])


def test_forward_without_gpu_raises(model):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    B, N = 1, 8
    p = torch.zeros(B, N, 3, 16, 16)
    pos = torch.zeros(B, N, 2)
    with torch.no_grad(), pytest.raises(Exception, match="CUDA|CPU"):
        model((p, p), (pos, pos), (None, None))


def test_gather_scores_gloo_world2():
    """N>1 path on CPU: two gloo ranks shard 7 pairs (ragged 4+3) and all-gather the scores in pair order."""
    script = os.path.join(ROOT, "tests", "_dist_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("DIST_OK") == 2, out.stdout + out.stderr


def test_gelu_polynomial_accuracy():
    """The fc1 epilogue's erf polynomial (csrc/gemm.cu: gelu_erf2), restated in fp32 numpy, against exact erf-GELU."""
    import re
    from scipy.special import erf
    src = open(os.path.join(ROOT, "vtamiq_b200", "csrc", "gemm.cu")).read()
    body = src[src.index("constexpr float kC[13]"):]
    coefs = [np.float32(v) for v in re.findall(r"(-?\d\.\d+e[+-]\d+)f", body[:body.index("};")])]
    assert len(coefs) == 13
    x = np.linspace(-8, 8, 400001).astype(np.float32)
    z = np.clip((x * np.float32(0.70710678118654752440)).astype(np.float32), np.float32(-3.5), np.float32(3.5))
    u = ((z * z).astype(np.float32) * np.float32(0.16326530612244897) - np.float32(1)).astype(np.float32)
    acc = np.full_like(x, coefs[12])
    for c in coefs[11::-1]:
        acc = (acc * u + c).astype(np.float32)
    e = (z * acc).astype(np.float32)
    h = (x * np.float32(0.5)).astype(np.float32)
    got = (h * e + h).astype(np.float32)
    want = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
    assert np.abs(got - want).max() < 4e-6
    # relative to fp16 resolution of the stored activation: well inside half an ulp wherever |gelu| >= 1e-2
    big = np.abs(want) >= 1e-2
    assert (np.abs(got - want)[big] / np.abs(want)[big]).max() < 2.5e-4


def test_layernorm_folding_algebra():
    """Engine.fold_layernorm: LN(x) W^T + b == rstd (x W'^T - mean colsum) + b' (what vtq_gemm_ln's epilogue applies),
    checked in fp64 on the fp16-rounded W' so that only the algebra is under test."""
    import torch
    from vtamiq_b200.engine import fold_layernorm
    g = torch.Generator().manual_seed(0)
    K, N, M, eps = 96, 40, 17, 1e-6
    x = torch.randn(M, K, generator=g, dtype=torch.float64) * 2 + 0.3
    w = torch.randn(N, K, generator=g) * 0.1
    b = torch.randn(N, generator=g)
    ln_w = torch.rand(K, generator=g) + 0.5
    ln_b = torch.randn(K, generator=g) * 0.2
    wf, bf, cs = fold_layernorm(w, b, ln_w, ln_b, torch.float16)
    assert wf.dtype == torch.float16 and bf.dtype == torch.float32 and cs.shape == (N,)
    mean = x.mean(1, keepdim=True)
    rstd = 1.0 / torch.sqrt(x.var(1, unbiased=False, keepdim=True) + eps)
    folded = rstd * (x @ wf.double().t() - mean * cs.double()[None, :]) + bf.double()[None, :]
    # same math with the LayerNorm applied explicitly, on the same rounded W' (divide the LN weight back out)
    ln = (x - mean) * rstd
    want = ln @ wf.double().t() + (b.double() + w.double() @ ln_b.double())[None, :]
    assert (folded - want).abs().max().item() < 1e-5
    # and against the textbook form with unrounded weights: only the fp16 rounding of W' apart
    ref = torch.nn.functional.layer_norm(x, (K,), ln_w.double(), ln_b.double(), eps) @ w.double().t() + b.double()
    assert (folded - ref).abs().max().item() < 2e-2


def test_empty_batch_and_shape_errors_need_no_device():
    """Edge cases handled on the host before any kernel is involved: B = 0 returns an empty score vector like the
    reference; mismatched ref/dist shapes and N = 0 raise."""
    import pytest
    import torch
    import vtamiq_b200
    m = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False, num_keep_layers=1)).eval()
    p0 = torch.zeros(0, 5, 3, 16, 16)
    with torch.no_grad():
        q, aux = m((p0, p0), (torch.zeros(0, 5, 2),) * 2, (None, None))
        assert q.shape == (0,) and q.dtype == torch.float32 and aux is None
        with pytest.raises(ValueError, match="same shape"):
            m((torch.zeros(1, 5, 3, 16, 16), torch.zeros(1, 4, 3, 16, 16)), (torch.zeros(1, 5, 2),) * 2, (None, None))
        with pytest.raises(ValueError, match="at least one patch"):
            m((torch.zeros(1, 0, 3, 16, 16),) * 2, (torch.zeros(1, 0, 2),) * 2, (None, None))


def test_grad_mode_never_returns_a_silently_detached_score():
    """With autograd recording, the module either has a backward for everything that wants gradients (the tail, with
    the encoder frozen) or raises — it never hands back a detached score.  Host-side decision only."""
    import pytest
    import torch
    import vtamiq_b200
    m = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False, num_keep_layers=1)).eval()
    p = torch.zeros(1, 4, 3, 16, 16)
    pos = torch.zeros(1, 4, 2)
    with pytest.raises(NotImplementedError, match="encoder parameters require grad"):
        m((p, p), (pos, pos), (None, None))          # eval mode, grad enabled, every parameter trainable
    for q in m.transformer.parameters():
        q.requires_grad = False
    assert m._grad_mode() == "tail"                  # frozen encoder: the differentiable tail takes over
    with pytest.raises(NotImplementedError, match="respect to the inputs"):
        m((p.clone().requires_grad_(True), p), (pos, pos), (None, None))
    for q in m.parameters():
        q.requires_grad = False
    assert m._grad_mode() == "none"
    with torch.no_grad():
        assert m._grad_mode() == "none"


def test_module_copies_and_pickles_without_its_engine():
    import copy
    import pickle
    import vtamiq_b200
    m = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False, num_keep_layers=1), operand_dtype="bf16").eval()
    c = copy.deepcopy(m)
    assert c.engine is not m.engine and c.engine.model is c and c.engine.operand_dtype == "bf16"
    r = pickle.loads(pickle.dumps(m))
    assert r.engine.model is r and r.engine.operand_dtype == "bf16"
    assert all((a == b).all() for a, b in zip(m.state_dict().values(), r.state_dict().values()))


def test_device_side_correlations_match_reference_and_scipy(golden_dir):
    """vtamiq_b200.metrics against utils/misc/correlations.py of the reference (golden fixture made by
    tests/golden/make_golden.py) and, for ties, against scipy directly."""
    import os
    import numpy as np
    import scipy.stats
    import torch
    from vtamiq_b200 import metrics as M
    z = np.load(os.path.join(golden_dir, "correlations.npz"))
    keys = ("SROCC", "KROCC", "PLCC", "RMSE", "PLCC_NOFIT", "RMSE_NOFIT")
    for k in range(3):
        a, b = torch.from_numpy(z[f"a{k}"]), torch.from_numpy(z[f"b{k}"])
        got = M.compute_correlations(a, b)
        want = dict(zip(keys, z[f"want{k}"]))
        for key in ("SROCC", "KROCC", "PLCC_NOFIT", "RMSE_NOFIT"):        # closed-form: to rounding
            assert abs(got[key] - want[key]) < 1e-10, (k, key, got[key], want[key])
        for key in ("PLCC", "RMSE"):                                       # behind an iterative least-squares fit
            assert abs(got[key] - want[key]) < 1e-6, (k, key, got[key], want[key])
        nofit = M.compute_correlations(a, b, fit=False)
        assert nofit["PLCC"] == nofit["PLCC_NOFIT"] and nofit["RMSE"] == nofit["RMSE_NOFIT"]
    rng = np.random.default_rng(5)
    a = np.round(rng.normal(size=500), 1)
    b = np.round(0.6 * a + rng.normal(size=500) * 0.7, 1)
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    assert np.abs(M.average_ranks(ta).numpy() - scipy.stats.rankdata(a)).max() == 0
    assert abs(float(M.spearman(ta, tb)) - scipy.stats.spearmanr(a, b).correlation) < 1e-12
    assert abs(float(M.kendall(ta, tb, block=96)) - scipy.stats.kendalltau(a, b).correlation) < 1e-12
    const = torch.full((7,), 3.0, dtype=torch.float64)
    assert torch.equal(M.normalize_array(const), torch.zeros(7, dtype=torch.float64))   # degenerate range: shifted only


def test_device_coordinate_sampler_has_the_reference_law(golden_dir):
    """vtamiq_b200.patch_sampling.perturbed_grid_samples vs raw draws of the reference's default
    sampler (tests/golden/sampler_draws.npz): same support, same one-sample-per-grid-cell structure, same jitter
    range, and marginals that a two-sample Kolmogorov-Smirnov test cannot tell apart."""
    import os
    import numpy as np
    import scipy.stats
    import torch
    from vtamiq_b200.patch_sampling import perturbed_grid_samples, sample_batch
    z = np.load(os.path.join(golden_dir, "sampler_draws.npz"))
    g = torch.Generator().manual_seed(3)
    for name in ("cfg2", "small", "tall"):
        ref = z[name].astype(np.float64)                       # (draws, 2, n)
        h, w, n = (int(v) for v in z[name + "_hwn"])
        ours = perturbed_grid_samples(ref.shape[0], h, w, 16, 16, n, device="cpu", generator=g).numpy()
        assert ours.shape == ref.shape and ours.dtype == np.float64
        width = max(int(np.ceil(np.sqrt(n / (h / w)))), 1)
        height = int(np.ceil(width * h / w))
        for smp in (ours, ref):
            assert smp[:, 0].min() >= 0 and smp[:, 0].max() <= h - 16 and smp[:, 1].min() >= 0 and smp[:, 1].max() <= w - 16
            cy = np.minimum(np.floor(smp[:, 0] / (h - 16) * height), height - 1)
            cx = np.minimum(np.floor(smp[:, 1] / (w - 16) * width), width - 1)
            cell = (cy * width + cx).astype(int)
            assert all(len(set(row)) == n for row in cell)     # n distinct grid cells per image
            off_y = smp[:, 0] / (h - 16) * height - cy - 0.5   # jitter inside the cell, |.| <= 2 * 0.2
            off_x = smp[:, 1] / (w - 16) * width - cx - 0.5
            assert np.abs(off_y).max() <= 0.4 + 1e-6 and np.abs(off_x).max() <= 0.4 + 1e-6
        for axis in (0, 1):
            p = scipy.stats.ks_2samp(ours[:, axis].ravel(), ref[:, axis].ravel()).pvalue
            assert p > 1e-3, (name, axis, p)
    # multi-scale budget: same level count / per-level counts as the host path, finest level first
    levels = sample_batch(3, 1024, 1024, 500, 16, 3, 2.0, device="cpu", generator=g)
    assert [t.shape for t in levels] == [(3, 2, 380), (3, 2, 96), (3, 2, 24)]
    assert float(levels[2][:, 0].max()) <= 256 - 16


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs first): stdout carries exactly one JSON line with the
    contract's keys; under torchrun every rank but 0 exits 0 without printing."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)), VTQ_CPU_BUDGET_S="4")
    other = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                           cwd=root, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), capture_output=True, text=True,
                           timeout=300)
    assert other.returncode == 0 and other.stdout.strip() == ""
    run = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1"], cwd=root,
                         env=env, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    lines = [ln for ln in run.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, run.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key
    from oracle import reference_runner
    want_kind = "reference" if reference_runner.available() else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the steps it reports are the steps it ran: value = pairs per step * steps / (steps * ms_per_step)
    assert d["steps"] == 2 and d["warmup"] == 1
    assert abs(d["value"] - d["pairs_per_cpu_step"] / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]
    # same workload description as the GPU arm prints for the same command line (the driver matches the two lines)
    import bench
    import argparse
    wl = bench.Workload(argparse.Namespace(config="cfg2", pairs=0, global_pairs=0, images="uint8"), 1, 0)
    assert d["config"] == wl.config and d["config"]["name"] == "cfg2" and d["config"]["pairs_per_step"] == 32


def test_bench_workloads_follow_baseline_configs():
    """bench.py --config: shapes of BASELINE.json's five configurations, weak and strong sharding."""
    import argparse
    import bench
    ns = lambda **k: argparse.Namespace(**{**dict(config="cfg2", pairs=0, global_pairs=0, images="uint8"), **k})
    shapes = {name: (bench.Workload(ns(config=name), 1, 0)) for name in bench.CONFIGS}
    assert (shapes["cfg1"].B, shapes["cfg1"].N, shapes["cfg1"].H, shapes["cfg1"].W) == (1, 256, 384, 512)
    assert (shapes["cfg2"].B, shapes["cfg2"].N) == (32, 500)
    assert (shapes["cfg3"].B, shapes["cfg3"].counts, shapes["cfg3"].H, shapes["cfg3"].vit) == (64, (380, 96, 24), 1024, {"num_scales": 3})
    assert (shapes["cfg4"].B, shapes["cfg4"].N, shapes["cfg4"].H, shapes["cfg4"].W, shapes["cfg4"].S) == (8, 5000, 2160, 3840, 5001)
    assert 256 <= shapes["cfg5"].B <= 2048
    weak = bench.Workload(ns(config="cfg5", pairs=256), 8, 3)
    assert (weak.B, weak.total_pairs, weak.scaling) == (256, 2048, "weak")
    strong = [bench.Workload(ns(config="cfg5", global_pairs=2050), 8, r) for r in range(8)]
    assert sum(w.B for w in strong) == 2050 and {w.B for w in strong} == {256, 257} and strong[0].scaling == "strong"
    fp = bench.flops_per_pair(500)
    assert abs(fp["total"] / 1e9 - 189.92) < 0.01 and abs(bench.flops_per_pair(5000)["total"] / 1e9 - 3554.8) < 0.1


def test_u8_transform_is_exact():
    """The vector gather kernels evaluate the reference's image transform — u/255 (to_tensor) then (t-.5)/.5
    (normalize), each rounded to fp32 — without divisions: q0 = u*r, e = fma(-q0,255,u), q = fma(e,r,q0),
    z = fma(q,2,-1) with r = RN(1/255) (csrc/gather.cu normalize_u8x2).  Exact rational arithmetic shows both forms
    round to the same fp32 value for every byte, and that value is what torch computes."""
    from fractions import Fraction
    import math

    def rn32(fr):
        if fr == 0:
            return Fraction(0)
        sign, a = (-1 if fr < 0 else 1), abs(fr)
        e = math.floor(math.log2(float(a)))
        while Fraction(2) ** e > a:
            e -= 1
        while Fraction(2) ** (e + 1) <= a:
            e += 1
        ulp = Fraction(2) ** (e - 23)
        qn = a / ulp
        n = qn.numerator // qn.denominator
        rem = qn - n
        if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and n % 2 == 1):
            n += 1
        return sign * n * ulp

    r = rn32(Fraction(1, 255))
    assert float(r) == float(np.float32(0.003921568859368563))
    t = torch.arange(256, dtype=torch.float32).div(255).sub_(0.5).div_(0.5).numpy()
    for u in range(256):
        U = Fraction(u)
        want_q = rn32(U / 255)
        q0 = rn32(U * r)
        e = U - q0 * 255
        assert rn32(e) == e                       # the fma residual is exactly representable
        q = rn32(q0 + e * r)
        assert q == want_q, u
        z = rn32(2 * q - 1)
        assert z == rn32(rn32(want_q - Fraction(1, 2)) * 2), u
        assert float(z) == float(t[u]), u
