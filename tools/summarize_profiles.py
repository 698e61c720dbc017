"""Turns the raw gpurun_out/ captures into the tracked summaries under profiles/ (launch-list shares, ncu raw page,
key-metric JSON of the dominant kernel, DRAM traffic read by bench.py).  Run here after a profiling call:

    python tools/summarize_profiles.py gpurun_out/launches_r1b.csv gpurun_out/prof_r1_final2.ncu-rep r01b
"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
agg = {}
for r in rows[-39:]:                       # the last 13 steps (3 launches each: list builder, fused, post pass): warm
    k = r[4]
    agg.setdefault(k, []).append(float(r[-1]) / 1e3)
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(ROOT, "profiles", "%s_launches_c2_step.csv" % tag), "w") as f:
    f.write("# ncu launch list of `python bench.py --profile-only --steps 2 --warmup 3` (C2 workload, device-resident loop), last 39 launches = 13 steps\n")
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none ; times are cold-cache and serialised: compare SHARES\n")
    f.write("kernel,launches,total_us,share_pct,avg_us\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write('"%s",%d,%.1f,%.1f,%.2f\n' % (k, len(v), sum(v), 100 * sum(v) / tot, sum(v) / len(v)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", "%s_ncu_fbank_fused_raw.csv" % tag), "w").write(raw)
rr = list(csv.reader(raw.splitlines()))
h, u, v = rr[0], rr[1], rr[2]
keep = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__icc_request_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]
d = {"note": sys.argv[4] if len(sys.argv) > 4 else ""}
for i, k in enumerate(h):
    if k in keep or ("issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k):
        d[k] = {"unit": u[i], "value": v[i]}
json.dump(d, open(os.path.join(ROOT, "profiles", "%s_ncu_fbank_fused.json" % tag), "w"), indent=1)
def val(k):
    x = d[k]; f_ = float(x["value"].replace(",", ""))
    return f_ * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[x["unit"]]
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
json.dump({"kernel": "fbank_fused_kernel<13,true,false,false>", "workload": "C2 ragged batch (seed 1), statistics mode",
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_step": rd + wr,
           "source": "profiles/%s_ncu_fbank_fused.json (ncu --set full, one launch = all fused work of one step)" % tag},
          open(os.path.join(ROOT, "profiles", "fbank_fused_summary.json"), "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", "%s_launches_c2_step.csv" % tag)).read())
print({k: d[k]["value"] for k in keep if k in d})
