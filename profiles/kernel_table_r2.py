"""Builds profiles/<tag>_kernels.md and profiles/ncu_traffic.json from the raw page of the `ncu --set full` capture of
scripts/ncu_forward.py (gpurun_out/<tag>_ncu_full_raw.csv, exported on the GPU box with `ncu -i ... --page raw --csv`):

    python profiles/kernel_table_r2.py r2j

Every captured launch is labelled from its kernel name and its position in the capture plan of scripts/ncu_forward.py
(ViT block 1 of the bench forward: 412 tokens in, k = 345 -> 347 out; the final norm; the image K / V^T projections;
text layer 0). Algorithmic bytes / FLOPs per launch follow DESIGN.md section 6 (operands read once, results written
once; error-compensation passes and recomputation are not credited)."""
import csv
import json
import re
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2j"
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6552.6}
HBM = peaks["hbm_gbs"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

# shapes of the captured launches (bench configuration 2, oracle trajectory k = 410, 345, ... -> 253)
B, H, d, T = 64, 12, 768, 100
N1, N2, NF, P = 412, 347, 255, 256          # tokens entering / leaving ViT block 1, final tokens, padded key rows
M1, M2 = B * N1, B * N2
MT, LT = 32 * 20, 20                        # text rows (32 sentences x 20 tokens)
MB = 1e6


def gemm(M, N, K, a_b, w_b, o_b, extra=0.0):
    return (M * K * a_b + N * K * w_b + M * N * o_b + extra) / MB, 2.0 * M * N * K


# label -> (kernel-name regex, description, algorithmic MB, algorithmic FLOP)
PLAN = [
    ("layernorm_kernel", "LN1 of ViT block 1 (+ fp16 hi/lo planes of y and x)", M1 * d * (4 + 8) / MB, 0),
    ("gemm_split3_kernel<128", "token . codebook^T, split fp16 planes (ViT block 1)", *gemm(M1, 128, d, 4, 4, 4)),
    ("token_colstats_kernel", "token_colstats (ViT block 1)", B * (N1 - 1) * 128 * 4 / MB, 0),
    ("query_sdft_mn_kernel", "query_sdft_planes: sd_ft from the x planes, MN-major operands (ViT block 1)",
     (B * (N1 - 1) * d * 4 + B * (N1 - 1) * 128 * 4 + 2 * B * T * d * 4) / MB, 2.0 * B * (N1 - 1) * T * d),
    ("gemm_split3_kernel<256", "q|k|v projection, split fp16 planes (ViT block 1)", *gemm(M1, 3 * d, d, 4, 4, 4)),
    ("attn_fwd_tc_kernel", "attention forward (ViT block 1)", (M1 * 3 * d * 4 + M1 * d * 2) / MB, 4.0 * B * H * N1 * N1 * 64),
    ("attn_stats_tc_kernel", "attention statistics (ViT block 1)", (M1 * 2 * d * 4) / MB, 2.0 * B * H * N1 * N1),
    ("attn_cls_combine_kernel", "CLS-row combine (until r2j; since fused into attn_stats_tc_kernel)", B * H * N1 * 4 * 2 / MB, 0),
    ("dtp_score_kernel", "dtp_score: cluster of 4 CTAs per sequence (ViT block 1)", B * (N1 - 1) * (T + 5) * 4 / MB, 0),
    ("gemm_tcgen05_kernel", "attention output projection + residual (ViT block 1)", *gemm(M1, d, d, 2, 2, 4, M1 * d * 4)),
    ("dtp_apply_kernel", "dtp_apply: radix select + gather + merge + norm2 (ViT block 1)", (M1 * d * 4 + M2 * d * 6) / MB, 0),
    ("gemm_tcgen05_kernel", "fc1 + GELU (ViT block 1)", *gemm(M2, 4 * d, d, 2, 2, 2)),
    ("gemm_tcgen05_kernel", "fc2 + residual (ViT block 1)", *gemm(M2, d, 4 * d, 2, 2, 4, M2 * d * 4)),
    ("layernorm_pack_kernel", "final norm -> fp16 cross-attention operand layout", (B * NF * d * 4 + B * NF * d * 2) / MB, 0),
    ("gemm_tcgen05_kernel", "image K projection, all 12 text layers (image0)", *gemm(32 * P, 12 * d, d, 2, 2, 2)),
    ("gemm_tcgen05_kernel", "image V^T projection, all 12 text layers (image0)", *gemm(12 * d, 32 * P, d, 2, 2, 2)),
    ("gemm_split3_kernel<128", "text q|k|v + codebook dots, split fp16 planes (layer 0)", *gemm(MT, 3 * d + 128, d, 4, 4, 4)),
    ("token_colstats_kernel", "token_colstats (text layer 0)", 32 * (LT - 1) * 128 * 4 / MB, 0),
    ("query_sdft_kernel", "query_sdft (text layer 0)", (32 * (LT - 1) * (d + 128) * 4 + 2 * 32 * T * d * 4) / MB, 0),
    ("small_self_attn_kernel", "text self-attention (layer 0)", (MT * 3 * d * 4 + MT * d * 2) / MB, 0),
    ("small_self_stats_kernel", "text self-attention statistics (layer 0)", 32 * H * LT * LT * 4 / MB, 0),
    ("dtp_score_kernel", "dtp_score (text layer 0)", 32 * (LT - 1) * (T + 5) * 4 / MB, 0),
    ("gemm_tcgen05_kernel", "self-output dense + residual (text layer 0)", *gemm(MT, d, d, 2, 2, 4, MT * d * 4)),
    ("dtp_apply_kernel", "dtp_apply (text layer 0)", (MT * d * 4 + MT * d * 6) / MB, 0),
    ("gemm_tcgen05_kernel", "twin cross-attention query projection (text layer 0)", *gemm(MT, 2 * d, d, 2, 2, 2)),
    ("cross_attn_tc_kernel", "cross-attention over 255 image tokens (text layer 0, branch 0)",
     (MT * d * 2 * 2 + 2 * 32 * P * d * 2) / MB, 4.0 * 32 * H * LT * NF * 64),
    ("gemm_tcgen05_kernel", "twin output dense, averaged (text layer 0)", *gemm(MT, d, 2 * d, 2, 2, 4, MT * d * 4)),
    ("gemm_tcgen05_kernel", "intermediate dense + GELU (text layer 0)", *gemm(MT, 4 * d, d, 2, 2, 2)),
    ("gemm_tcgen05_kernel", "output dense + residual (text layer 0)", *gemm(MT, d, 4 * d, 2, 2, 4, MT * d * 4)),
]


def load(path):
    rows = list(csv.reader(open(path)))
    head, unit = rows[0], dict(zip(rows[0], rows[1]))
    return [dict(zip(head, r)) for r in rows[2:]], unit


def val(dd, unit, key):
    try:
        return float(dd[key].replace(",", "")) * SCALE.get(unit.get(key, ""), 1.0)
    except Exception:
        return float("nan")


src = ROOT / "gpurun_out" / f"{TAG}_ncu_full_raw.csv"
rows, unit = load(src)
shutil.copyfile(src, ROOT / "profiles" / f"{TAG}_ncu_full_raw.csv")
out = [f"# Per-kernel ncu evidence -- {TAG} (final kernels of round 2)", "",
       "`ncu --set full --clock-control none --profile-from-start off` on `scripts/ncu_forward.py`: one BLIP-NLVR forward at the",
       "bench configuration (32 pairs = 64 images 384x384, temperature 3.5894, device-resident lengths, Python-issued launches =",
       "the launches the CUDA graph replays), profiler switched on around the launches of ViT block 1 (412 tokens in, k = 345),",
       f"the final norm, the image K / V^T projections and text layer 0. Raw export: `profiles/{TAG}_ncu_full_raw.csv`.",
       "Durations are ncu-serialised and cold-cache (compare shares, not absolutes).",
       f"Denominators (MEASURED_PEAKS.json): HBM {HBM} GB/s; tensor % = sm__pipe_tensor_cycles_active, pct of peak sustained active.",
       "`alg MB` = algorithmic bytes of the launch (operands read once + results written once); `traffic/alg` = measured DRAM",
       "bytes over that (< 1: part of the operands was still in the 126 MB L2).", "",
       "| # | launch | kernel | grid | time us | tensor % | DRAM MB (r + w) | DRAM GB/s | % HBM peak | alg MB | traffic/alg | alg TFLOP/s | regs |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {"all": {}}
pi = 0
for i, dd in enumerate(rows):
    k = re.sub(r"\(.*", "", dd["Kernel Name"]).replace("void ", "").replace("madtp::", "")
    label, alg_mb, alg_fl = "(other launch inside the capture window)", float("nan"), 0.0
    for j in range(pi, len(PLAN)):
        if k.startswith(PLAN[j][0]):
            label, alg_mb, alg_fl = PLAN[j][1], PLAN[j][2], PLAN[j][3]
            pi = j + 1
            break
    if k.startswith("at::") or "elementwise" in k:
        k = "torch " + k.split("::")[-1][:36]
    t = val(dd, unit, "gpu__time_duration.sum")
    rd, wr = val(dd, unit, "dram__bytes_read.sum"), val(dd, unit, "dram__bytes_write.sum")
    gbs = (rd + wr) / (t * 1e-6) / 1e9 if t > 0 else float("nan")
    tens = val(dd, {}, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    tf = alg_fl / (t * 1e-6) / 1e12 if (t > 0 and alg_fl) else float("nan")
    ratio = (rd + wr) / (alg_mb * MB) if alg_mb == alg_mb and alg_mb > 0 else float("nan")
    out.append(f"| {i} | {label} | {k} | {dd.get('launch__grid_size', '')} | {t:.1f} | {tens:.1f} | {rd / 1e6:.1f} + {wr / 1e6:.1f} | "
               f"{gbs:.0f} | {100 * gbs / HBM:.1f} | {alg_mb:.1f} | {ratio:.2f} | {tf:.0f} | {dd.get('launch__registers_per_thread', '')} |")
    traffic["all"][f"{i}:{label}"] = {"kernel": k, "time_us": t, "dram_bytes_per_launch": rd + wr,
                                      "algorithmic_bytes": None if alg_mb != alg_mb else int(alg_mb * MB), "tensor_pct": tens}
(ROOT / "profiles" / f"{TAG}_kernels.md").write_text("\n".join(out) + "\n")
# the dominant class of the bench line (madtp_gemm:f16): its largest member, fc1 + GELU of ViT block 1
f = {k: v for k, v in traffic["all"].items()}
pick = lambda s: next((v for k, v in f.items() if s in k), None)   # noqa: E731
fc1, fc2, proj = pick("fc1 + GELU"), pick("fc2 + residual"), pick("attention output projection")
if fc1:
    traffic["madtp_gemm:f16"] = {
        "dram_bytes_per_launch": fc1["dram_bytes_per_launch"], "algorithmic_bytes": fc1["algorithmic_bytes"],
        "traffic_over_algorithmic": fc1["dram_bytes_per_launch"] / fc1["algorithmic_bytes"],
        "launch": f"fc1 + GELU of ViT block 1 (M = {M2}, N = 3072, K = 768), the largest member of the class; other members: "
                  f"fc2 {fc2['dram_bytes_per_launch'] / 1e6:.1f} MB measured / {fc2['algorithmic_bytes'] / 1e6:.1f} MB algorithmic, "
                  f"proj {proj['dram_bytes_per_launch'] / 1e6:.1f} / {proj['algorithmic_bytes'] / 1e6:.1f}",
        "source": f"profiles/{TAG}_kernels.md (ncu --set full, scripts/ncu_forward.py)"}
(ROOT / "profiles" / "ncu_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
print("\n".join(out[12:]))
