"""Timeline of attention CTA 0 (clock64 stamps recorded by the kernel itself) — run on the GPU box."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vtamiq_b200 import _lib
ctx = _lib.get_context(0)
n_seq, S, heads = 64, 501, 12
H = heads * 64
qkv = (torch.randn(n_seq * S, 3 * H, device="cuda") * 1.5).half()
out = torch.empty(n_seq * S, H, device="cuda", dtype=torch.float16)
tr = torch.zeros(3, 512, dtype=torch.int64, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    tr.zero_()
    ctx.call("vtq_attention_fwd_trace", P(qkv), P(out), n_seq, S, heads, 0, P(tr), st)
torch.cuda.synchronize()
t = tr.cpu().numpy()
t0 = t[t > 0].min()
mma = t[0][t[0] > 0] - t0
print("MMA thread stamps (first 40):", mma[:40].tolist())
if os.environ.get("VTQ_ATTN_V3") == "1" or os.environ.get("VTQ_ATTN_V5") == "1":
    sys.exit("attn_trace.py decodes the stamps of the current kernel only")
# current kernel: role 1 = softmax warp (tile A, key half 0), 4 stamps per key tile (tile start, S in registers,
# exponentials start, P published); role 2 = helper warp, 4 stamps per key tile (wait S_A, max_A published, wait S_B,
# max_B published) — the epilogue of the previous work item sits between two tiles' stamps
nkv = (S + 127) // 128
r = t[1][t[1] > 0] - t0
print(f"softmax (A, half 0): {len(r)} stamps, total span {r[-1] - r[0]} cycles, {(r[-1] - r[0]) / (len(r) / 4):.0f} cycles per key tile")
for i in range(0, min(len(r) - 4, 4 * 4 * nkv), 4):
    seg = r[i:i + 5]
    tile = i // 4
    print(f"  item {tile // nkv} tile {tile % nkv}: start {int(seg[0]):7d} | wait S + ld {int(seg[1]-seg[0]):5d} | mask/max/exchange {int(seg[2]-seg[1]):5d} | exp {int(seg[3]-seg[2]):5d} | to next tile {int(seg[4]-seg[3]):5d}")
