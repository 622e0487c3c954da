"""Same-shape library reference for the encoder projections: torch.matmul (cuBLASLt fp16, what the reference's
autocast forward runs, train.py:602 + transformer.py:154-156,:169,:213-214) beside vtq_gemm on the benchmark shapes.
Each is timed alone with CUDA events over back-to-back launches on inputs larger than L2 (rotating buffers), so both
see the same burst clocks.  Prints one JSON line per shape; run under gpurun, copy into profiles/."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vtamiq_b200 import _lib  # noqa: E402

P = lambda t: None if t is None else C.c_void_p(t.data_ptr())


def timeit(fn, iters):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    M = 2 * pairs * 501
    ctx = _lib.get_context(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    shapes = [("gemm_qkv", 2304, 768, 0), ("gemm_out", 768, 768, 3), ("gemm_fc1", 3072, 768, 1), ("gemm_fc2", 768, 3072, 3)]
    only = os.environ.get("GVC_SHAPES")          # e.g. GVC_SHAPES=gemm_out,gemm_fc2
    if only:
        shapes = [s for s in shapes if s[0] in only.split(",")]
    vtq_only = os.environ.get("GVC_VTQ_ONLY") == "1"
    nbuf = 4   # rotate operand / output buffers so that nothing is served from L2 across launches
    for name, N, K, epi in shapes:
        g = torch.Generator(device="cuda").manual_seed(N + K)
        A = [(torch.randn(M, K, device="cuda", generator=g) * 0.5).half() for _ in range(nbuf)]
        W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
        b = torch.randn(N, device="cuda", generator=g)
        out32 = epi == 3
        O = [torch.zeros(M, N, device="cuda", dtype=torch.float32 if out32 else torch.float16) for _ in range(nbuf)]

        def ours(i):
            ctx.call("vtq_gemm", P(A[i % nbuf]), 0, P(W), P(b), M, N, K, 0, epi, P(O[i % nbuf]), 0, None, st)

        Wt = W.t().contiguous()       # cuBLAS picks its own preferred layout either way; give it both a try
        C16 = [torch.empty(M, N, device="cuda", dtype=torch.float16) for _ in range(nbuf)]

        def cublas_nt(i):
            torch.matmul(A[i % nbuf], W.t(), out=C16[i % nbuf])

        def cublas_nn(i):
            torch.matmul(A[i % nbuf], Wt, out=C16[i % nbuf])

        def cublas_linear(i):   # what nn.Linear runs under autocast: addmm with bias epilogue
            torch.nn.functional.linear(A[i % nbuf], W, b.half())

        iters = 50
        t_ours = timeit(ours, iters)
        if vtq_only:
            print(json.dumps({"shape": name, "M": M, "N": N, "K": K, "bn_n768": os.environ.get("VTQ_GEMM_BN_N768"),
                              "vtq_gemm": {"ms": round(t_ours, 4), "tflops": round(2.0 * M * N * K / (t_ours * 1e-3) / 1e12, 1)}}))
            continue
        t_nt, t_nn, t_lin = timeit(cublas_nt, iters), timeit(cublas_nn, iters), timeit(cublas_linear, iters)
        fl = 2.0 * M * N * K
        tf = lambda ms: round(fl / (ms * 1e-3) / 1e12, 1)
        print(json.dumps({"shape": name, "M": M, "N": N, "K": K,
                          "vtq_gemm": {"ms": round(t_ours, 4), "tflops": tf(t_ours),
                                       "epilogue": ["bias->fp16", "bias+GELU->fp16", "", "bias + fp32 residual add"][epi]},
                          "cublas_matmul_nt": {"ms": round(t_nt, 4), "tflops": tf(t_nt)},
                          "cublas_matmul_nn": {"ms": round(t_nn, 4), "tflops": tf(t_nn)},
                          "cublas_linear_bias": {"ms": round(t_lin, 4), "tflops": tf(t_lin)},
                          "note": "cuBLAS rows are plain fp16 GEMMs (no GELU / residual work); standalone launches"}))


if __name__ == "__main__":
    main()
