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
names = ["waitS", "Sready", "Sregs", "max", "pvdone", "turn", "Ppub"]
for role in (1, 2):
    r = t[role][t[role] > 0] - t0
    print(f"softmax {'AB'[role-1]}: {len(r)} stamps, total span {r[-1] - r[0]} cycles")
    # per tile: 6 stamps, +1 per item (4 tiles per item at S=501)
    i = 0
    item = 0
    while i + 7 <= len(r) and item < 3:
        for j in range(4):
            seg = r[i:i + 7]
            if len(seg) < 7: break
            d = [int(seg[k + 1] - seg[k]) for k in range(6)]
            print(f"  item {item} tile {j}: start {int(seg[0]):7d} | wait S {d[0]:5d} | ld S {d[1]:5d} | max {d[2]:5d} | wait PV+rescale {d[3]:5d} | wait turn {d[4]:5d} | exp+P {d[5]:5d}")
            i += 7
        if i < len(r):
            print(f"  item {item} output: {int(r[i] - r[i-1])} cycles")
            i += 1
        item += 1

if os.environ.get("ATT_TIMELINE"):   # ATT_TIMELINE=1 [ATT_LO=.. ATT_HI=..]: merged event list of CTA 0
    # merged timeline of one work item (the second): MMA issues vs softmax phases
    ev = []
    m = t[0][t[0] > 0] - t0
    # MMA stamps per item at S=501: 2 (first QK A,B) + 4 tiles * (up to 4) ... just label sequentially
    for i, x in enumerate(m): ev.append((int(x), f"MMA issue #{i}"))
    names = ["start(wait S)", "S ready", "S in regs", "max done", "PV retired", "turn acquired", "P published"]
    for role in (1, 2):
        r = t[role][t[role] > 0] - t0
        i = 0; item = 0
        while i < len(r):
            for j in range(4):
                for k in range(7):
                    if i < len(r): ev.append((int(r[i]), f"{'AB'[role-1]} item{item} tile{j} {names[k]}")); i += 1
            if i < len(r): ev.append((int(r[i]), f"{'AB'[role-1]} item{item} output issued")); i += 1
            item += 1
    ev.sort()
    lo, hi = int(os.environ.get("ATT_LO", 18000)), int(os.environ.get("ATT_HI", 36000))
    prev = None
    for x, s in ev:
        if lo <= x <= hi:
            print(f"{x:8d} (+{0 if prev is None else x - prev:5d}) {s}")
            prev = x
