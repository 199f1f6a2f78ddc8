"""Temperature calibration fixtures for BASELINE configurations 3-5 -- TEST INFRASTRUCTURE (run in the build container).

    python -m oracle.gen_calib retrieval|vqa|clip [batch]

For every configuration the ORACLE (oracle/dtp_oracle.py, pinned to the reference by oracle/gen_golden.py) runs the
seeded synthetic batch at a bisected temperature until MACs(pruned) / MACs(unpruned) = 1 - p (closed form of the pruning
trajectory, madtp_b200/flops.py), and the fixture stores the temperature, the trajectories, the per-layer keep-masks
of the image tower and the outputs, so that bench.py can report the GPU run's agreement without a CPU-heavy pass.
  retrieval  BASELINE config 3: BLIP retrieval evaluation path (compress_retrieval_dtp.py:104,120,170-177), 384 x 384,
             text padded to 35, p = 0.75
  vqa        BASELINE config 5: BLIP-VQA image + question encoders (models/blip_vqa.py:60,119-125), 480 x 480, p = 0.5
  clip       BASELINE config 4: CLIP ViT-B/16 towers (clip/model.py:482-503), 336 x 336, context 77, p = 0.5
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from madtp_b200 import flops, synthetic  # noqa: E402
from oracle import dtp_oracle as O  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def ks_of(traces):
    return [t.k if t.pruned else -1 for t in traces]


def bisect(run, p, lo=0.25, hi=128.0, tol=0.004, iters=14):
    best = None
    for it in range(iters):
        mid = (lo * hi) ** 0.5
        t0 = time.time()
        r, payload = run(mid)
        print(f"it{it}: T={mid:.4f} ratio={r:.4f} ({time.time() - t0:.1f}s)", flush=True)
        if best is None or abs(r - (1 - p)) < abs(best[1] - (1 - p)):
            best = (mid, r, payload)
        if abs(r - (1 - p)) < tol:
            break
        if r > 1 - p:
            lo = mid
        else:
            hi = mid
    return best


def keeps(out, prefix, traces):
    for i, t in enumerate(traces):
        if t.pruned:
            out[f"{prefix}{i}_keep"] = np.packbits(t.keep.numpy(), axis=1)
            out[f"{prefix}{i}_count"] = t.count.numpy()


def gen_retrieval(batch=64, p=0.75, size=384, max_len=35):
    sd = synthetic.retrieval_state_dict(4321, img_size=size)
    images, ids, mask = synthetic.retrieval_inputs(batch, size, max_len, seed=0)
    space = sd["space_dict"]
    n0 = (size // 16) ** 2 + 1
    full = flops.retrieval_macs(n0, [-1] * 12, max_len, [-1] * 12, [-1] * 12)
    ids2 = ids.clone()
    ids2[:, 0] = 30523

    def run(temp):
        tv, tt, tm = [], [], []
        with torch.no_grad():
            feat, _ = O.vit_forward(images, sd, "visual_encoder.", space, temp, traces=tv)
            txt, _ = O.med_text_encoder(ids, mask, sd, "text_encoder.", None, space, temp, "text", traces=tt)
            mm, _ = O.med_text_encoder(ids2, mask, sd, "text_encoder.", feat, space, temp, "multimodal", traces=tm)
            itm = O.linear(mm[:, 0, :], sd, "itm_head")
        macs = flops.retrieval_macs(n0, ks_of(tv), max_len, ks_of(tt), ks_of(tm))
        return macs / full, (tv, tt, tm, feat, txt, itm, macs)
    temp, r, (tv, tt, tm, feat, txt, itm, macs) = bisect(run, p)
    assert abs(r - (1 - p)) < 0.01, r
    out = {"temperature": np.array(temp), "ratio": np.array(r), "batch": np.array(batch), "p": np.array(p),
           "image_size": np.array(size), "text_len": np.array(max_len),
           "input_digest": np.array(synthetic.tensor_digest(images, ids, mask)),
           "vit_k": np.array(ks_of(tv)), "text_k": np.array(ks_of(tt)), "mm_k": np.array(ks_of(tm)),
           "itm": itm.numpy(), "text_cls": txt[:, 0, :].numpy(), "image_cls": feat[:, 0, :].numpy(),
           "macs_pruned": np.array(macs), "macs_unpruned": np.array(full)}
    keeps(out, "vit", tv)
    np.savez_compressed(GOLDEN / f"calib_retrieval_p{int(p * 100)}_b{batch}.npz", **out)
    print("retrieval:", temp, r, ks_of(tv), ks_of(tt), ks_of(tm))


def gen_vqa(batch=64, p=0.5, size=480, max_len=20):
    sd = synthetic.vqa_state_dict(99, img_size=size)
    images, ids, mask = synthetic.retrieval_inputs(batch, size, max_len, seed=2)     # padding='longest' -> 19 tokens max
    longest = int(mask.sum(1).max())
    ids, mask = ids[:, :longest].contiguous(), mask[:, :longest].contiguous()
    space = sd["space_dict"]
    n0 = (size // 16) ** 2 + 1
    full = flops.vqa_encoder_macs(n0, [-1] * 12, longest, [-1] * 12)
    ids2 = ids.clone()
    ids2[:, 0] = 30523

    def run(temp):
        tv, tm = [], []
        with torch.no_grad():
            feat, _ = O.vit_forward(images, sd, "visual_encoder.", space, temp, traces=tv)
            q, _ = O.med_text_encoder(ids2, mask, sd, "text_encoder.", feat, space, temp, "multimodal", traces=tm)
        macs = flops.vqa_encoder_macs(n0, ks_of(tv), longest, ks_of(tm))
        return macs / full, (tv, tm, feat, q, macs)
    temp, r, (tv, tm, feat, q, macs) = bisect(run, p)
    assert abs(r - (1 - p)) < 0.01, r
    out = {"temperature": np.array(temp), "ratio": np.array(r), "batch": np.array(batch), "p": np.array(p),
           "image_size": np.array(size), "text_len": np.array(longest),
           "input_digest": np.array(synthetic.tensor_digest(images, ids, mask)),
           "vit_k": np.array(ks_of(tv)), "mm_k": np.array(ks_of(tm)),
           "question_cls": q[:, 0, :].numpy(), "image_cls": feat[:, 0, :].numpy(),
           "macs_pruned": np.array(macs), "macs_unpruned": np.array(full)}
    keeps(out, "vit", tv)
    np.savez_compressed(GOLDEN / f"calib_vqa_p{int(p * 100)}_b{batch}.npz", **out)
    print("vqa:", temp, r, ks_of(tv), ks_of(tm))


def gen_clip(batch=64, p=0.5, size=336):
    sd = synthetic.clip_state_dict(777, img_size=size)
    images, text = synthetic.clip_inputs(batch, size, seed=0)
    space = sd["space_dict"]
    n0 = (size // 16) ** 2 + 1
    full = flops.clip_macs(n0, [-1] * 12, [-1] * 12)

    def run(temp):
        tv, tt = [], []
        with torch.no_grad():
            img, _ = O.clip_vision_forward(images, sd, "visual.", space, temp, 12, 12, traces=tv)
            txt, _ = O.clip_text_forward(text, sd, space, temp, 12, 8, traces=tt)
        macs = flops.clip_macs(n0, ks_of(tv), ks_of(tt))
        return macs / full, (tv, tt, img, txt, macs)
    temp, r, (tv, tt, img, txt, macs) = bisect(run, p)
    assert abs(r - (1 - p)) < 0.01, r
    out = {"temperature": np.array(temp), "ratio": np.array(r), "batch": np.array(batch), "p": np.array(p),
           "image_size": np.array(size), "input_digest": np.array(synthetic.tensor_digest(images, text)),
           "vision_k": np.array(ks_of(tv)), "text_k": np.array(ks_of(tt)), "image_emb": img.numpy(),
           "text_emb": txt.numpy(), "macs_pruned": np.array(macs), "macs_unpruned": np.array(full)}
    keeps(out, "vit", tv)
    np.savez_compressed(GOLDEN / f"calib_clip_p{int(p * 100)}_b{batch}_r{size}.npz", **out)
    print("clip:", temp, r, ks_of(tv), ks_of(tt))


if __name__ == "__main__":
    what = sys.argv[1]
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    torch.set_num_threads(4)
    {"retrieval": gen_retrieval, "vqa": gen_vqa, "clip": gen_clip}[what](batch)
