"""Builds profiles/r1e_kernels.md from the raw ncu pages exported on the GPU box (gpurun_out/r1e_{vit,text}_raw.csv)."""
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


out = ["# Per-kernel ncu evidence -- round 1, final kernels", "",
       "One BLIP-NLVR forward at the bench configuration (32 pairs = 64 images 384x384, temperature 3.5894): every launch",
       "of two consecutive pruned ViT layers (about 347 and 321 tokens x 64 images) and of one text layer (20 tokens x 32",
       "sentences, 255 image tokens each), captured with `ncu --set full --clock-control none` on `scripts/layer_once.py`",
       "and exported on the GPU box (`--page raw --csv`); each window is a contiguous slice of the launch stream.",
       "Durations are ncu-serialised and cold-cache (compare shares, not absolutes).",
       f"Denominators (MEASURED_PEAKS.json): HBM {HBM} GB/s; tensor % = sm__pipe_tensor_cycles_active, pct of peak sustained active.",
       "", "| kernel | grid | time us | tensor % | DRAM MB (r + w) | DRAM GB/s | % of HBM peak | regs |", "|---|---|---|---|---|---|---|---|"]
for title, name in (("two consecutive ViT layers", "r1e_vit_raw.csv"), ("one NLVR text layer", "r1e_text_raw.csv")):
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
(ROOT / "profiles" / "r1e_kernels.md").write_text("\n".join(out) + "\n")
print("\n".join(out[:60]))
