"""profiles/<round>_summary.md from the bench lines + ncu summaries that scripts/make_profiles.sh copied."""
import json
import re
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
d = json.load(open(f"profiles/{R}_bench.json"))
ref = json.load(open(f"profiles/{R}_bench_reference.json"))
rf, pk = d["roofline"], d["frac_of_bf16_peak"]


def ncu_metric(path, kernel_pat, metric_pat):
    """first value of a metric in an ncu_summary.py text dump for the first kernel matching kernel_pat"""
    try:
        txt = open(path).read()
    except OSError:
        return None
    for block in txt.split("\n\n"):
        if re.search(kernel_pat, block):
            m = re.search(metric_pat + r"[^\n]*?([0-9][0-9.,]*)\s*$", block, re.M)
            if m:
                return m.group(1)
    return None


out = []
w = out.append
w(f"# Round {int(R[1:])} — measured summary (B200, {d['config']['workload']})\n")
w(f"* own arm: **{d['value']} pairs/s** ({d['ms_per_step']} ms/step, CUDA-graph replay, inputs resident in HBM); "
  f"end-to-end through `VTAMIQ.forward` from pinned host patches: **{d['e2e']['value']} pairs/s** "
  f"({d['e2e']['h2d_bytes_per_step'] / 1e6:.1f} MB H2D per step).")
w(f"* reference arm ({ref['cpu_baseline']['kind']} of the reference's CPU forward, {ref['cpu_baseline']['cores']} host "
  f"cores): {ref['value']:.2f} pairs/s.")
w(f"* algorithmic work {d['algorithmic_gflop_per_pair']} GFLOP/pair -> {d['achieved_tflops_algorithmic']} TFLOP/s = "
  f"{pk['burst']} of measured bf16 burst, {pk['sustained']} of sustained.")
w(f"* roofline object ({rf['kernel']}): {rf['achieved']} {rf['unit']} executed, frac {rf['frac']} of {rf['peak']} "
  f"(= {rf['frac_of_burst']} of burst, {rf['frac_of_sustained']} of sustained; {rf['peak_src']}); share of step {rf['share_of_step']}; DRAM traffic per launch "
  f"{rf['traffic'] / 1e6:.0f} MB vs algorithmic {rf['algorithmic_bytes_per_launch'] / 1e6:.0f} MB.")
w(f"* clocks during the timed region: {d['clocks']}")
w(f"* launches per step: {d['launches_per_step']}\n")
w("| launch class | launches/step | avg ms (CUDA events, un-graphed) | share | TFLOP/s |")
w("|---|---|---|---|---|")
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["share_of_step"]):
    w(f"| {k} | {v['launches_per_step']} | {v['avg_ms']} | {v['share_of_step']} | {v.get('tflops', '')} |")
w("")
w(f"ncu evidence: `{R}_launches.csv` / `{R}_launches.md` (launch list, `gpu__time_duration.sum`), `{R}_ncu_gemm2.txt`, "
  f"`{R}_ncu_attention.txt`, `{R}_ncu_hbm.txt` (`--set full` summaries), `gemm_traffic.json` (DRAM bytes per launch).")
try:
    w("")
    w(open(f"profiles/{R}_notes.md").read().rstrip())
except OSError:
    pass
open(f"profiles/{R}_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
