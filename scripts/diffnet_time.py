"""Stand-alone time of vtq_diffnet_head (default VTAMIQ tail: 4 RG x 4 RCAB, hidden 768) at B pairs."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vtamiq_b200
from vtamiq_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
m = vtamiq_b200.VTAMIQ(vit_config=dict(pretrained=False)).eval().to("cuda:0")
eng = m.engine

N = 64
patches = torch.randn(2, max(B, 1), N, 3, 16, 16, device="cuda")
pos = torch.rand(2, max(B, 1), N, 2, device="cuda")
with torch.no_grad():
    m((patches[0], patches[1]), (pos[0], pos[1]), None)   # builds the packed parameter list
ctx = eng.ctx
hidden = 768
diff = torch.randn(B, hidden, device="cuda")
q = torch.empty(B, device="cuda")
ws = torch.empty(ctx.workspace_bytes(B, hidden), dtype=torch.uint8, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def call():
    ctx.call("vtq_diffnet_head", P(diff), eng._tail_params, len(eng._tail_params), eng.num_rgs, eng.num_rcabs, hidden,
             eng.ca_hidden, eng.head_hidden, B, P(q), P(ws), st)
for _ in range(3): call()
torch.cuda.synchronize()
for cold in (False, True):
    ts = []
    for _ in range(20):
        if cold: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(" ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("VTQ_")), f"B={B} diffnet ms ({'L2 flushed' if cold else 'weights in L2'}) median {ts[len(ts)//2]:.4f} min {ts[0]:.4f}")
