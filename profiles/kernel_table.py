"""Builds profiles/<tag>_kernels.md from the raw ncu page exported on the GPU box (gpurun_out/<tag>_layers_raw.csv):
    python profiles/kernel_table.py r1f"""
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
HBM = peaks["hbm_gbs"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def load(path):
    rows = list(csv.reader(open(path)))
    head, unit = rows[0], dict(zip(rows[0], rows[1]))
    return [dict(zip(head, r)) for r in rows[2:]], unit


def val(d, unit, key):
    try:
        return float(d[key].replace(",", "")) * SCALE.get(unit.get(key, ""), 1.0)
    except Exception:
        return float("nan")


TAG = sys.argv[1] if len(sys.argv) > 1 else "r1f"
out = [f"# Per-kernel ncu evidence -- {TAG}", "",
       "One BLIP-NLVR forward at the bench configuration (32 pairs = 64 images 384x384, temperature 3.5894): every launch",
       "of ViT blocks 2 and 3 (about 347 and 321 tokens x 64 images) and of text layer 1 (about 20 tokens x 32 sentences,",
       "255 image tokens each), captured with `ncu --set full --clock-control none --nvtx --nvtx-include cap/` on",
       "`scripts/layer_once.py` and exported on the GPU box (`--page raw --csv`).",
       "Durations are ncu-serialised and cold-cache (compare shares, not absolutes).",
       f"Denominators (MEASURED_PEAKS.json): HBM {HBM} GB/s; tensor % = sm__pipe_tensor_cycles_active, pct of peak sustained active.",
       "", "| kernel | grid | time us | tensor % | DRAM MB (r + w) | DRAM GB/s | % of HBM peak | regs |", "|---|---|---|---|---|---|---|---|"]
for title, name in (("ViT blocks 2, 3 and text layer 1, in launch order", f"{TAG}_layers_raw.csv"),):
    rows, unit = load(ROOT / "gpurun_out" / name)
    out.append(f"| **{title}** | | | | | | | |")
    for d in rows:
        k = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("madtp::", "")
        if k.startswith("at::"):
            k = "torch " + k.split("::")[-1][:32]
        t = val(d, unit, "gpu__time_duration.sum")
        rd, wr = val(d, unit, "dram__bytes_read.sum"), val(d, unit, "dram__bytes_write.sum")
        gbs = (rd + wr) / (t * 1e-6) / 1e9 if t > 0 else float("nan")
        tens = val(d, {}, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        out.append(f"| {k} | {d.get('launch__grid_size', '')} | {t:.1f} | {tens:.1f} | {rd / 1e6:.1f} + {wr / 1e6:.1f} | "
                   f"{gbs:.0f} | {100 * gbs / HBM:.1f} | {d.get('launch__registers_per_thread', '')} |")
(ROOT / "profiles" / f"{TAG}_kernels.md").write_text("\n".join(out) + "\n")
print("\n".join(out[:60]))
