"""Round-2 profile summaries: gpurun_out/r03_launches.csv + r03_prof_fused.ncu-rep + r03_prof_post.ncu-rep -> profiles/r03_*.
Also writes profiles/r03_step_traffic.json (DRAM bytes per launch of the two kernels of a C2 step + the hash of the CUDA sources
the capture profiled), which bench.py reads for `roofline.traffic`.

    python tools/summarize_r03.py [note]
"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import source_hash
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
note = sys.argv[1] if len(sys.argv) > 1 else ""
rows = [r for r in csv.reader(open(os.path.join(G, "r03_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
agg = {}
warm = rows[-21:]                         # the last 7 steps (3 launches each: list builder, fused, post pass): warm
for r in warm:
    agg.setdefault(r[4].split("(")[0], []).append(float(r[-1]) / 1e3)
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(P, "r03_launches_c2_step.csv"), "w") as f:
    f.write("# ncu launch list of `python bench.py --profile-only --steps 2 --warmup 3` (C2 workload, device-resident loop), last 21 launches = 7 steps\n")
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none ; times are cold-cache and serialised: compare SHARES\n")
    f.write("kernel,launches,total_us,share_pct,avg_us\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write('"%s",%d,%.1f,%.1f,%.2f\n' % (k, len(v), sum(v), 100 * sum(v) / tot, sum(v) / len(v)))
keep = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "memory_l1_wavefronts_shared",
        "memory_l1_wavefronts_shared_ideal", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__icc_request_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]


def summarize(rep, out_json, out_raw, extra):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(out_raw, "w").write(raw)
    rr = list(csv.reader(raw.splitlines()))
    h, u, v = rr[0], rr[1], rr[2]
    d = dict(extra)
    for i, k in enumerate(h):
        if k in keep or ("issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k):
            d[k] = {"unit": u[i], "value": v[i]}
    json.dump(d, open(out_json, "w"), indent=1)
    return d


def val(d, k):
    x = d[k]
    return float(x["value"].replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3}.get(x["unit"], 1.0)


frames = 460361          # frames of the C2 batch (seed 1)
fu = summarize(os.path.join(G, "r03_prof_fused.ncu-rep"), os.path.join(P, "r03_ncu_fbank_fused.json"), os.path.join(P, "r03_ncu_fbank_fused_raw.csv"),
               {"note": "round 2: fused kernel (lean instantiation, statistics mode, padding tiles) on the C2 ragged batch; power spectra inside the transposition "
                        "buffers + split-twiddle rotation; ncu --set full --clock-control none -k regex:fbank_fused -s 5 -c 1 of `bench.py --profile-only`. " + note})
po = summarize(os.path.join(G, "r03_prof_post.ncu-rep"), os.path.join(P, "r03_ncu_postpass.json"), os.path.join(P, "r03_ncu_postpass_raw.csv"),
               {"note": "round 2: in-place utterance-CMVN post pass of the same C2 step. " + note})
cyc = val(fu, "sm__cycles_elapsed.avg")
wf = val(fu, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
derived = {"frames": frames, "cycles_per_frame_per_sm": cyc / (frames / 148.0),
           "shared_wavefronts_per_frame": wf / frames, "warp_instructions_per_frame": val(fu, "smsp__inst_executed.sum") / frames,
           "r02_for_comparison": {"cycles_per_frame_per_sm": 607988.5 / (frames / 148.0), "shared_wavefronts_per_frame": 54164206 / frames,
                                  "warp_instructions_per_frame": 186973815 / frames, "registers": 124, "smem_kB_per_cta": 104.3, "warps_active_pct": 24.84,
                                  "gpu_time_us": 322.8}}
fu["derived"] = derived
json.dump(fu, open(os.path.join(P, "r03_ncu_fbank_fused.json"), "w"), indent=1)
traffic = {"source_hash": source_hash(), "workload": "C2 ragged batch (seed 1), utterance CMVN: list builder + fused launch + post pass per step",
           "fused_dram_bytes_per_launch": val(fu, "dram__bytes_read.sum") + val(fu, "dram__bytes_write.sum"),
           "fused_dram_bytes_read": val(fu, "dram__bytes_read.sum"), "fused_dram_bytes_write": val(fu, "dram__bytes_write.sum"),
           "postpass_dram_bytes_per_launch": val(po, "dram__bytes_read.sum") + val(po, "dram__bytes_write.sum"),
           "postpass_dram_bytes_read": val(po, "dram__bytes_read.sum"), "postpass_dram_bytes_write": val(po, "dram__bytes_write.sum"),
           "captures": ["profiles/r03_ncu_fbank_fused.json", "profiles/r03_ncu_postpass.json"]}
traffic["whole_step_dram_bytes"] = traffic["fused_dram_bytes_per_launch"] + traffic["postpass_dram_bytes_per_launch"]
json.dump(traffic, open(os.path.join(P, "r03_step_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, "r03_launches_c2_step.csv")).read())
print(json.dumps(derived, indent=1))
print(json.dumps(traffic, indent=1))
