"""DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per GEMM launch from the ncu --set full capture
(gpurun_out/prof_gemm2.ncu-rep) -> profiles/gemm_traffic.json, which bench.py reports as roofline.traffic."""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
def val(d, k):
    v = float(d[k]); u = units[hdr.index(k)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
caps = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    caps.append((d["Kernel Name"], float(d["gpu__time_duration.sum"]), val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum")))
res = {}
n192 = sorted([c for c in caps if "192, 1, 0" in c[0] or "256, 1, 0" in c[0]], key=lambda c: c[1])   # fp32 residual epilogue
if len(n192) >= 2:
    res["gemm_out"], res["gemm_fc2"] = n192[0][2], n192[-1][2]
for name, _, b in caps:
    if "256, 0, 1" in name: res["gemm_fc1"] = b
    if "256, 0, 0" in name: res["gemm_qkv"] = b
res["source"] = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, cfg2 shapes (M=32064)"
json.dump(res, open(out, "w"), indent=1)
print(res)
