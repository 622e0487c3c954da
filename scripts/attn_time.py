"""Stand-alone kernel time of vtq_attention_fwd at the benchmark shape (64 sequences x 501 tokens x 12 heads)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vtamiq_b200 import _lib
ctx = _lib.get_context(0)
n_seq, S, heads = 64, 501, 12
H = heads * 64
qkv = torch.randn(n_seq * S, 3 * H, device="cuda").half()
out = torch.empty(n_seq * S, H, device="cuda", dtype=torch.float16)
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(5):
    ctx.call("vtq_attention_fwd", P(qkv), P(out), n_seq, S, heads, 0, 0, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    ctx.call("vtq_attention_fwd", P(qkv), P(out), n_seq, S, heads, 0, 0, st)
e1.record(); torch.cuda.synchronize()
print(" ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("VTQ_")), "attention ms", round(e0.elapsed_time(e1) / 100, 4))
