#!/usr/bin/env python
"""Benchmark of the VTAMIQ inference hot path on B200 — BASELINE.json metric: ref/dist pairs/sec, ViT-B/16.

    python bench.py --gpus 1 --steps 30 --warmup 5                 # cfg2 (the configuration the metric is quoted on)
    python bench.py --config cfg4                                   # any of BASELINE.json's configs: cfg1 .. cfg5
    python bench.py --config cfg5 --global-pairs 2048 --gpus N      # strong scaling: ONE global batch split over N GPUs
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # CPU arm: the reference itself (oracle/_ref) or its oracle port

One "step" = one pass of the hot path over one batch: device patch gather (decoded uint8 images + sampled
coordinates resident in HBM; the reference's image transform is fused into the gather) -> patch embedding -> 12
encoder blocks for ref and dist -> CLS difference -> DiffNet -> scores; the scores of all ranks are gathered once at
the end of the timed region.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

HIDDEN, MLP, LAYERS = 768, 3072, 12
METRIC = "ref/dist pairs/sec (VTAMIQ ViT-B/16 forward)"

# BASELINE.json configs: pairs per GPU per step, image size (H, W), patches per image per scale
# (finest first), extra ViT kwargs.  cfg5 is the throughput sweep (256-2048 pairs; --pairs / --global-pairs).
CONFIGS = {
    "cfg1": dict(pairs=1, hw=(384, 512), counts=(256,), vit={},
                 text="cfg1: 1 pair 512x384, 256 single-scale 16x16 patches per image (latency case)"),
    "cfg2": dict(pairs=32, hw=(384, 512), counts=(500,), vit={},
                 text="cfg2: batch 32 pairs 512x384, 500 single-scale 16x16 patches per image"),
    "cfg3": dict(pairs=64, hw=(1024, 1024), counts=(380, 96, 24), vit=dict(num_scales=3),
                 text="cfg3: batch 64 pairs 1024x1024, 500 patches over 3 scales (380/96/24 at 16/32/64 px), "
                      "scale embeddings"),
    "cfg4": dict(pairs=8, hw=(2160, 3840), counts=(5000,), vit={},
                 text="cfg4: batch 8 pairs 3840x2160, 5000 single-scale patches per image (S = 5001)"),
    "cfg5": dict(pairs=1024, hw=(384, 512), counts=(500,), vit={},
                 text="cfg5: throughput sweep point, 512x384, 500 single-scale patches per image"),
}


def flops_per_pair(n_patches: int, tokens: int = 1) -> dict:
    """Algorithmic FLOPs: 1 MAC = 2 FLOP, full dense math of the reference."""
    S = n_patches + tokens
    embed = 2 * n_patches * HIDDEN * HIDDEN
    linear = LAYERS * S * 2 * (4 * HIDDEN * HIDDEN + 2 * HIDDEN * MLP)
    attn = LAYERS * 4 * S * S * HIDDEN
    tail = 29.79e6
    return dict(embed=2 * embed, linear=2 * linear, attn=2 * attn, tail=tail,
                total=2 * (embed + linear + attn) + tail)


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(src="measured", hbm=d["hbm_gbs"], burst=d["bf16_tflops"], sustained=d["bf16_tflops_sustained"])
    return dict(src="fallback", hbm=6650.0, burst=1590.0, sustained=1400.0)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index: int, period_s: float = 0.01):
        super().__init__(daemon=True)
        self.index, self.period, self.stop_flag = index, period_s, threading.Event()
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is None:
            return
        n = self.nvml
        names = {getattr(n, k): k[len("nvmlClocksEventReason"):] for k in dir(n) if k.startswith("nvmlClocksEventReason")
                 and isinstance(getattr(n, k), int)}
        names.update({getattr(n, k): k[len("nvmlClocksThrottleReason"):] for k in dir(n)
                      if k.startswith("nvmlClocksThrottleReason") and isinstance(getattr(n, k), int)
                      and getattr(n, k) not in names})
        while not self.stop_flag.is_set():
            try:
                self.sm.append(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if bit and (mask & bit) == bit and nm not in ("None", "All", "GpuIdle"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self) -> dict:
        self.stop_flag.set()
        if self.nvml is None or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        import re
        snake = lambda nm: re.sub(r"(?<!^)(?=[A-Z])", "_", nm).lower()
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(snake(r) for r in self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------------ workload
class Workload:
    """One BASELINE config resolved against the command line: shapes, per-rank batch, descriptive `config` dict
    (identical for both arms so the driver can match them)."""

    def __init__(self, args, world: int, rank: int):
        from vtamiq_b200.parallel import shard_pairs
        c = CONFIGS[args.config]
        self.name = args.config
        self.H, self.W = c["hw"]
        self.counts = tuple(c["counts"])
        self.N = int(sum(self.counts))
        self.vit = dict(c["vit"])
        self.S = self.N + 1
        if args.global_pairs:
            a, b = shard_pairs(args.global_pairs, rank, world)
            self.B, self.total_pairs, self.scaling = b - a, args.global_pairs, "strong"
            self.offset = a
        else:
            self.B = args.pairs if args.pairs else c["pairs"]
            self.total_pairs, self.scaling, self.offset = self.B * world, "weak", rank * self.B
        if self.B < 1:
            raise SystemExit("bench.py: every rank needs at least one pair")
        text = c["text"]
        if self.name == "cfg5" or args.pairs or args.global_pairs:
            text += f" [{self.total_pairs} pairs per step over {world} GPU(s)]"
        self.config = {
            "workload": text + ", ViT-B/16 + DiffNet", "name": self.name,
            "pairs_per_step": self.total_pairs, "patches": self.N, "patches_per_scale": list(self.counts),
            "image_hw": [self.H, self.W],
            "images": ("synthetic uint8 HWC images (as decoded)" if args.images == "uint8"
                       else "synthetic fp32 CHW images (already normalised)") + " + float64 patch coordinates",
            "parallelism": f"dp{world} (pairs sharded, weight replicas, one score all_gather per run)",
            "l2": "no explicit flush: the per-step working set (activations + patch matrix + images, two alternating "
                  "input sets) is far larger than the 126 MB L2",
        }


def synth_inputs(wl: Workload, B: int, seed: int, images_kind: str):
    """images: uint8 (2,B,H,W,3) or fp32 (2,B,3,H,W); samples: list over scales of float64 (B,2,n_s).
    Images cycle over a small pool of distinct synthetic pairs (content does not change the work, generation
    time does)."""
    import synth
    H, W = wl.H, wl.W
    pool = 2 if H * W > 2_000_000 else 4
    pool = min(pool, B)
    rng = np.random.default_rng(seed)
    levels = synth.graded_levels(max(pool, 2), seed)
    base = []
    for p in range(pool):
        # white-noise reference family for the very large images (the 1/f family needs an FFT of the whole frame)
        ref, dist = synth.make_pair(seed * 100 + p, H, W, float(levels[p]), family="white" if H * W > 2_000_000 else "auto")
        base.append(torch.from_numpy(np.stack([ref, dist])))            # (2,H,W,3) uint8
    u8 = torch.stack([base[b % pool] for b in range(B)], dim=1).contiguous()   # (2,B,H,W,3)
    samples = []
    for s, n in enumerate(wl.counts):
        samples.append(torch.from_numpy(np.stack([synth.jittered_samples(rng, H >> s, W >> s, n) for _ in range(B)])))
    if images_kind == "uint8":
        return u8, samples
    f32 = u8.permute(0, 1, 4, 2, 3).contiguous().to(torch.float32).div_(255).sub_(0.5).div_(0.5)
    return f32, samples


def build_model(vit_cfg=None, device=None, dtype="fp16"):
    import synth
    import vtamiq_b200
    torch.manual_seed(0)
    m = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False, **(vit_cfg or {})), operand_dtype=dtype).eval()
    synth.perturb_(m)
    return m.to(device) if device is not None else m


# ------------------------------------------------------------------------------------------------ CPU arm
class CpuReference:
    """The reference's CPU implementation of the path on all host cores: the UNMODIFIED reference when the build
    container packed it (oracle/_ref, kind "reference"), else the oracle port (same ATen CPU ops, kind "port").
    One pair = the reference's per-pair host gather (get_iqa_patches) + its share of a batched fp32 forward."""

    def __init__(self, wl: Workload):
        import synth
        from oracle import reference_runner
        self.wl = wl
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.kind = "reference" if reference_runner.available() else "port"
        if self.kind == "reference":
            self.runner = reference_runner
            self.model = reference_runner.build_model(vit_cfg=wl.vit, perturb=synth.perturb_)
        else:
            self.sd = build_model(wl.vit).state_dict()
        self.u8, self.samples = synth_inputs(wl, min(wl.B, 64), seed=7, images_kind="uint8")

    def step(self, pairs: int):
        """gather + forward of `pairs` pairs; returns the scores."""
        import synth
        from oracle import patch_oracle, vtamiq_oracle
        wl = self.wl
        P, POS, SC = [], [], []
        pool = self.u8.shape[1]
        for i in range(pairs):
            p = i % pool
            ref, dist = self.u8[0, p].numpy(), self.u8[1, p].numpy()
            tens = (synth.to_tensor_normalized(ref), synth.to_tensor_normalized(dist))   # the reference's transform
            smp = [s[p].numpy() for s in self.samples]
            if self.kind == "reference":
                pp, pos, sc = self.runner.gather(ref, dist, tens, smp, len(wl.counts))
            else:
                pp, pos, sc = patch_oracle.extract_patches(np.stack([t.numpy() for t in tens]), smp)
                pp, pos = torch.from_numpy(pp), torch.from_numpy(pos)
                sc = None if sc is None else torch.from_numpy(sc)
            P.append(pp); POS.append(pos); SC.append(sc)
        P, POS = torch.stack(P), torch.stack(POS)
        use_sc = SC[0] is not None
        SCt = torch.stack(SC).to(torch.float32) if use_sc else None       # train.py:254 casts the ids to fp32
        c = lambda t: t.contiguous()           # train.py:258-267 hands the model cloned (contiguous) slices
        scales = (c(SCt[:, 0]), c(SCt[:, 1])) if use_sc else (None, None)
        with torch.no_grad():
            if self.kind == "reference":
                q, _ = self.model((c(P[:, 0]), c(P[:, 1])), (c(POS[:, 0]), c(POS[:, 1])), scales)
                return q
            return vtamiq_oracle.vtamiq_forward(self.sd, (c(P[:, 0]), c(P[:, 1])), (c(POS[:, 0]), c(POS[:, 1])),
                                                scales if use_sc else None)

    def sample_rate(self, pairs: int, repeats: int):
        best = float("inf")
        for _ in range(repeats):
            t0 = time.perf_counter()
            self.step(pairs)
            best = min(best, time.perf_counter() - t0)
        what = "the unmodified reference (oracle/_ref)" if self.kind == "reference" else "oracle port of the reference"
        return pairs / best, (f"{pairs} pair(s) x {self.wl.N} patches, {self.wl.H}x{self.wl.W}: per-pair host gather + "
                              f"batched fp32 forward of {what}, best of {repeats}")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wl = Workload(args, world, 0)
    t_start = time.perf_counter()
    cpu = CpuReference(wl)
    # calibrate on one pair (also warms the allocator / thread pool), then size the per-step sample so that the
    # whole --steps K --warmup W run stays within ~2.5 minutes
    t0 = time.perf_counter()
    cpu.step(1)
    t1 = time.perf_counter() - t0
    budget_s = float(os.environ.get("VTQ_CPU_BUDGET_S", "150"))
    n_steps = max(args.steps, 1) + max(args.warmup, 0)
    pairs = int(max(1, min(wl.B, budget_s / n_steps / max(t1, 1e-3))))
    for _ in range(args.warmup):
        cpu.step(pairs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.step(pairs)
    dt = time.perf_counter() - t0
    rate = pairs * args.steps / dt
    what = "the unmodified reference (oracle/_ref)" if cpu.kind == "reference" else "oracle port of the reference"
    sample = (f"{what}: each step = {pairs} of the configuration's {wl.B} pairs per GPU x {wl.N} patches "
              f"({wl.H}x{wl.W}), per-pair host gather + one batched fp32 forward, {cpu.cores} threads; "
              f"{args.steps} timed steps after {args.warmup} warm-up steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": wl.config, "pairs_per_cpu_step": pairs,
        "cpu_baseline": {"value": rate, "unit": "pairs/s", "cores": cpu.cores, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": rate, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_start,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def kernel_flops(wl: Workload, B: int) -> dict:
    """Executed FLOPs per launch of every tensor-core launch class (2B sequences per launch)."""
    S, N = wl.S, wl.N
    rows, n_seq = 2 * B * S, 2 * B
    return {
        "gemm_embed": 2.0 * (n_seq * N) * HIDDEN * HIDDEN,
        "gemm_qkv": 2.0 * rows * HIDDEN * 3 * HIDDEN,
        "gemm_out": 2.0 * rows * HIDDEN * HIDDEN,
        "gemm_fc1": 2.0 * rows * HIDDEN * MLP,
        "gemm_fc2": 2.0 * rows * HIDDEN * MLP,
        "attention": 4.0 * n_seq * S * S * HIDDEN,
        # last block, quality-token rows only (2B rows); attention: one 256-row granule against all S keys
        "gemm_out_tok": 2.0 * n_seq * HIDDEN * HIDDEN,
        "gemm_fc1_tok": 2.0 * n_seq * HIDDEN * MLP,
        "gemm_fc2_tok": 2.0 * n_seq * HIDDEN * MLP,
        "attention_tok": 4.0 * n_seq * min(256, S) * S * HIDDEN,
    }


def kernel_bytes(wl: Workload, B: int, images_kind: str) -> dict:
    """Algorithmic HBM bytes per launch of the bandwidth-bound launch classes."""
    S, N = wl.S, wl.N
    rows, n_seq = 2 * B * S, 2 * B
    px = 1 if images_kind == "uint8" else 4
    return {
        "patch_gather": n_seq * N * 768 * px + n_seq * N * 768 * 2,      # all levels together (per step, not per launch)
        "embed_assemble": n_seq * N * 768 * (4 + (4 if len(wl.counts) > 1 else 0)) + rows * 768 * 4,
        "layernorm": rows * 768 * (4 + 2),
    }


def profile_leg(model, eng, sets, nprof: int):
    """Per-launch CUDA events over `nprof` instrumented (un-graphed) steps -> {tag: [ms, ...]}."""
    eng.timeline = []
    for i in range(nprof):
        model.forward_from_images(*sets[i & 1], validate="off")
    torch.cuda.synchronize()
    agg = {}
    for tag, a, b in eng.timeline:
        agg.setdefault(tag, []).append(a.elapsed_time(b))
    eng.timeline = None
    return agg


def gemm_rate(agg, fl, nprof):
    tags = [t for t in ("gemm_qkv", "gemm_out", "gemm_fc1", "gemm_fc2") if t in agg]
    ms = sum(float(np.sum(agg[t])) for t in tags) / nprof
    flops = sum(fl[t] * (len(agg[t]) // nprof) for t in tags)
    return tags, ms, flops, (flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0)


def run_gpu_arm(args):
    import torch.distributed as dist
    from vtamiq_b200 import extract_patches_batch
    from vtamiq_b200.parallel import gather_scores

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    torch.set_grad_enabled(False)     # inference: the call context of train.py:592-602 (model.eval() + no_grad)
    dev = torch.device("cuda", local)
    wl = Workload(args, world, rank)
    B = wl.B
    model = build_model(wl.vit, dev, args.dtype)
    eng = model.engine

    # two distinct input sets, alternated, resident in HBM
    sets = []
    for s in range(2):
        images, samples = synth_inputs(wl, B, seed=1 + 2 * rank + s, images_kind=args.images)
        sets.append((images.to(dev), [t.to(dev) for t in samples]))
    total_pairs = wl.total_pairs

    # Pairs are independent: the ranks run their batches with NO per-step collective; every step's scores
    # stay in a device buffer and ONE all_gather at the end of the timed region hands all of them to every rank.
    longest = -(-total_pairs // world)
    score_buf = torch.zeros(max(args.steps, 1), longest, dtype=torch.float32, device=dev)

    def step(i):
        images, samples = sets[i & 1]
        q = model.forward_from_images(images, samples, validate="off")
        score_buf[i % score_buf.shape[0], :B].copy_(q)
        return q

    def collect(n_steps):
        flat = score_buf[:n_steps].reshape(-1)
        return gather_scores(flat, flat.numel() * world) if world > 1 else flat

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # launch accounting: one un-graphed step, counted by the library itself; then capture the CUDA graph
    # (both before NCCL comes up, so no communicator thread is alive during stream capture)
    eng.use_cuda_graph = False
    model.forward_from_images(*sets[0], validate="sync")     # coordinates validated once, outside the timed region
    model.forward_from_images(*sets[1], validate="sync")
    torch.cuda.synchronize()
    n0 = eng.ctx.launch_count()
    model.forward_from_images(*sets[1], validate="off")
    torch.cuda.synchronize()
    launches_per_step = eng.ctx.launch_count() - n0
    eng.use_cuda_graph = not args.no_graph
    model.forward_from_images(*sets[0], validate="off")
    torch.cuda.synchronize()
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    warm = max(args.warmup, 3)
    for i in range(warm):
        step(i)
    collect(min(warm, args.steps))   # the collective is warm too
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(i)
    all_scores = collect(args.steps)
    ev1.record()
    barrier()
    assert all_scores.numel() == args.steps * longest * world
    clocks = sampler.result()
    ms_local = ev0.elapsed_time(ev1)
    per_rank_ms = [ms_local]
    if world > 1:
        t = torch.tensor([ms_local], device=dev, dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [float(x.item()) for x in allt]
    ms = max(per_rank_ms)
    value = total_pairs * args.steps / (ms / 1e3)

    # ---- e2e (reference-facing API): VTAMIQ.forward(patches, pos, scales) from HOST (pinned) fp32 patches; H2D of the
    # step's inputs and D2H of the scores inside the timed region; uploads double-buffered on a copy stream.
    # ---- e2e_from_images: the repo's fast entry from pinned uint8 images + coordinates (forward_from_images).
    copy_stream = torch.cuda.Stream()

    def e2e_run(host_sets, call, n_warm, n_steps):
        dev_bufs = [[torch.empty_like(t, device=dev) for t in host_sets[0]] for _ in range(2)]
        q_host = torch.empty(B, dtype=torch.float32).pin_memory()
        done_compute = [None, None]

        def upload(i):
            with torch.cuda.stream(copy_stream):
                for d, h in zip(dev_bufs[i & 1], host_sets[i & 1]):
                    d.copy_(h, non_blocking=True)
                e = torch.cuda.Event()
                e.record(copy_stream)
            return e

        def loop(n):
            ready = upload(0)
            for i in range(n):
                nxt = None
                if i + 1 < n:
                    if done_compute[(i + 1) & 1] is not None:
                        copy_stream.wait_event(done_compute[(i + 1) & 1])   # buffer reuse: its last reader finished
                    nxt = upload(i + 1)
                torch.cuda.current_stream().wait_event(ready)
                with torch.no_grad():
                    qd = call(dev_bufs[i & 1])
                ev = torch.cuda.Event()
                ev.record()
                done_compute[i & 1] = ev
                q_host.copy_(qd, non_blocking=True)   # this rank's scores; ranks exchange nothing per step
                ready = nxt
            torch.cuda.synchronize()

        loop(n_warm)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop(n_steps)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d = sum(t_.numel() * t_.element_size() for t_ in host_sets[0])
        return total_pairs * n_steps / (float(t.item()) / 1e3), h2d, wall

    n_levels = len(wl.counts)
    use_sc = eng.scale_table is not None
    host_ref, host_img = [], []
    with torch.no_grad():
        for s in range(2):
            im, sm = sets[s]
            P, POS, SC = extract_patches_batch(im, sm, with_scales=use_sc)
            hs = [P.cpu().pin_memory(), POS.cpu().pin_memory()] + ([SC.cpu().pin_memory()] if use_sc else [])
            host_ref.append(hs)
            host_img.append([im.cpu().pin_memory()] + [t.cpu().pin_memory() for t in sm])
            del P, POS, SC
    torch.cuda.empty_cache()

    def call_reference_api(bufs):
        P, POS = bufs[0], bufs[1]
        sc = (bufs[2][0], bufs[2][1]) if use_sc else (None, None)
        return model((P[0], P[1]), (POS[0], POS[1]), sc)[0]

    def call_from_images(bufs):
        return model.forward_from_images(bufs[0], list(bufs[1:1 + n_levels]), validate="off")

    e2e_value, h2d_bytes, e2e_wall = e2e_run(host_ref, call_reference_api, warm, args.steps)
    e2i_value, h2d_img_bytes, e2i_wall = e2e_run(host_img, call_from_images, warm, args.steps)
    del host_ref, host_img
    d2h_bytes = B * 4

    # ---- roofline leg (rank 0): per-launch CUDA events over instrumented (un-graphed) steps
    roofline, breakdown, sustained = None, None, None
    if rank == 0:
        pk = peaks()
        fl = kernel_flops(wl, B)
        by = kernel_bytes(wl, B, args.images)
        nprof = 3
        agg = profile_leg(model, eng, sets, nprof)
        breakdown = {}
        step_ms = sum(sum(v) for v in agg.values()) / nprof
        gather_ms = sum(float(np.sum(agg[t])) for t in ("patch_gather",) if t in agg) / nprof
        for tag, v in sorted(agg.items()):
            avg = float(np.mean(v))
            d = {"launches_per_step": len(v) // nprof, "avg_ms": round(avg, 4),
                 "share_of_step": round(sum(v) / nprof / step_ms, 4)}
            if tag in fl:
                d["tflops"] = round(fl[tag] / (avg * 1e-3) / 1e12, 1)
                d["frac_of_burst"] = round(d["tflops"] / pk["burst"], 4)
            if tag in by:
                per_launch = by[tag] / (1 if tag != "patch_gather" else 1)
                t_ms = gather_ms if tag == "patch_gather" else avg
                d["algorithmic_GBps"] = round(per_launch / (t_ms * 1e-3) / 1e9, 1)
                d["frac_of_hbm"] = round(d["algorithmic_GBps"] / pk["hbm"], 4)
            breakdown[tag] = d
        gemm_tags, g_ms, g_fl, achieved = gemm_rate(agg, fl, nprof)
        traffic = None   # DRAM bytes per launch (average over the same launches), from the committed ncu capture
        tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tpath) and wl.name == "cfg2" and B == CONFIGS["cfg2"]["pairs"]:
            with open(tpath) as fh:
                tj = json.load(fh)
            if all(t_ in tj for t_ in gemm_tags):
                n_l = sum(len(agg[t_]) // nprof for t_ in gemm_tags)
                traffic = round(sum(tj[t_] * (len(agg[t_]) // nprof) for t_ in gemm_tags) / n_l)
        region_s = ms / 1e3
        short = region_s < 1.0
        peak = pk["burst"] if short else pk["sustained"]
        rows = 2 * B * wl.S
        n_gl = max(sum(len(agg[t_]) // nprof for t_ in gemm_tags), 1)
        roofline = {
            "bound": "tensor", "kernel": "vtq::gemm2_kernel (encoder QKV/out/fc1/fc2 projections, tcgen05 cta_group::2)",
            "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
            "frac_of_burst": round(achieved / pk["burst"], 4), "frac_of_sustained": round(achieved / pk["sustained"], 4),
            "peak_src": f"MEASURED_PEAKS.json ({pk['src']}): bf16 burst {pk['burst']}, sustained {pk['sustained']}; "
                        f"timed region {region_s:.2f} s -> {'burst' if short else 'sustained'} denominator",
            "flops_per_step": g_fl, "avg_ms_per_step": round(g_ms, 3), "share_of_step": round(g_ms / step_ms, 4),
            "note": "executed FLOPs of the full-row projection launches, CUDA events per launch over 3 un-graphed "
                    "steps (the last block's token-row launches are listed separately under kernels as *_tok)",
            "traffic": traffic,
            "algorithmic_bytes_per_launch": round(sum(
                {"gemm_qkv": rows * HIDDEN * 2 + rows * 3 * HIDDEN * 2, "gemm_out": rows * HIDDEN * 2 + 2 * rows * HIDDEN * 4,
                 "gemm_fc1": rows * HIDDEN * 2 + rows * MLP * 2, "gemm_fc2": rows * MLP * 2 + 2 * rows * HIDDEN * 4}[t_]
                * (len(agg[t_]) // nprof) for t_ in gemm_tags) / n_gl),
        }
        # ---- sustained leg: >= 3 s of back-to-back steps (the power cap's steady state), then the per-launch events
        # again while the board is hot.  Local to rank 0 (no collective inside).
        if not args.no_sustained:
            target_s = 3.0
            n_sus = int(max(10, min(20000, target_s / max(ms / 1e3 / max(args.steps, 1), 1e-5))))
            torch.cuda.synchronize()
            smp2 = ClockSampler(local)
            smp2.start()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(n_sus):
                model.forward_from_images(*sets[i & 1], validate="off")
            s1.record()
            torch.cuda.synchronize()
            sus_clocks = smp2.result()
            sus_ms = s0.elapsed_time(s1)
            agg2 = profile_leg(model, eng, sets, nprof)
            _, g_ms2, g_fl2, ach2 = gemm_rate(agg2, fl, nprof)
            fp_ = flops_per_pair(wl.N)
            sus_value = B * n_sus / (sus_ms / 1e3)
            sustained = {
                "value": round(sus_value, 2), "unit": "pairs/s (this GPU)", "steps": n_sus,
                "region_s": round(sus_ms / 1e3, 3), "ms_per_step": round(sus_ms / n_sus, 4), "clocks": sus_clocks,
                "achieved_tflops_algorithmic": round(sus_value * fp_["total"] / 1e12, 1),
                "frac_of_sustained_peak_algorithmic": round(sus_value * fp_["total"] / 1e12 / pk["sustained"], 4),
                "gemm": {"achieved": round(ach2, 1), "peak": pk["sustained"], "unit": "TFLOP/s",
                         "frac": round(ach2 / pk["sustained"], 4),
                         "note": "encoder projection launches, per-launch events right after the sustained loop"},
            }

    # ---- CPU baseline (rank 0, N=1 only): bounded sample on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ref = CpuReference(wl)
        n_s = 1 if wl.N > 1000 else min(B, 8)
        rate, sample = ref.sample_rate(n_s, repeats=2 if wl.N > 1000 else 3)
        cpu = {"value": round(rate, 3), "unit": "pairs/s", "cores": ref.cores, "kind": ref.kind, "sample": sample}

    if rank == 0:
        fp = flops_per_pair(wl.N)
        pk = peaks()
        per_gpu = value / world
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": wl.scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": wl.config,
            "implementation": {"operands": f"{args.dtype} tcgen05 operands, fp32 accumulate/residual/LN/softmax/DiffNet",
                               "cuda_graph": not args.no_graph, "pairs_per_gpu": B},
            "per_rank_ms": [round(x, 3) for x in per_rank_ms],
            "straggler_rank": int(np.argmax(per_rank_ms)),
            "algorithmic_gflop_per_pair": round(fp["total"] / 1e9, 2),
            "achieved_tflops_algorithmic": round(per_gpu * fp["total"] / 1e12, 1),
            "frac_of_bf16_peak": {"burst": round(per_gpu * fp["total"] / 1e12 / pk["burst"], 4),
                                  "sustained": round(per_gpu * fp["total"] / 1e12 / pk["sustained"], 4)},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes,
                    "api": "VTAMIQ.forward(patches, pos, scales) from pinned host fp32 patches; uploads double-buffered",
                    "wall_s": round(e2e_wall, 3)},
            "e2e_from_images": {"value": round(e2i_value, 2), "unit": "pairs/s", "h2d_bytes_per_step": h2d_img_bytes,
                                "d2h_bytes_per_step": d2h_bytes,
                                "api": f"VTAMIQ.forward_from_images(images, samples) from pinned host {args.images} "
                                       "images + float64 coordinates (gather inside the timed region)",
                                "wall_s": round(e2i_wall, 3)},
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "roofline": roofline, "sustained": sustained, "kernels": breakdown, "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_RESULT_OUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write there too (NCCL prints its version banner to
    fd 1).  Keep a private handle on the real stdout for the result and point fd 1 / sys.stdout at stderr."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line: dict):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS),
                    help="BASELINE.json configuration (default cfg2: the one the metric is quoted on)")
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--images", default="uint8", choices=["uint8", "fp32"],
                    help="resident image format: uint8 HWC as decoded (default) or fp32 CHW already normalised")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 3 s sustained leg")
    ap.add_argument("--pairs", type=int, default=0,
                    help="pairs per GPU per step (weak scaling; default: the configuration's batch)")
    ap.add_argument("--global-pairs", type=int, default=0,
                    help="strong scaling: ONE global batch of this many pairs split evenly over the GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
