"""Turns the scratch ncu outputs under gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r1_launches_step.csv
    python profiles/summarize.py full gpurun_out/prof_attn_r1.ncu-rep profiles/r1_ncu_attn.csv
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
           "smsp__inst_executed.sum"]


def launches(src, dst):
    with open(src) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    starts = [i for i, r in enumerate(rows) if "patchify" in r["Kernel Name"]]
    step = rows[starts[-2]:starts[-1]] if len(starts) >= 2 else rows
    agg = collections.OrderedDict()
    for r in step:
        nm = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        d = agg.setdefault(nm, [0, 0.0])
        d[0] += 1
        d[1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches_per_step", "us_per_step", "share_pct"])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, v[0], f"{v[1]:.1f}", f"{100 * v[1] / tot:.1f}"])
        w.writerow(["TOTAL (one BLIP-NLVR forward, 32 pairs; ncu-serialised, cold cache)", len(step), f"{tot:.1f}", "100.0"])


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(m, hdr.index(m)) for m in ["Kernel Name"] + METRICS if m in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([m for m, _ in cols])
        w.writerow([units[i] for _, i in cols])
        for r in rows[2:]:
            w.writerow([re.sub(r"\(.*", "", r[i]) if m == "Kernel Name" else r[i] for m, i in cols])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
