"""One BLIP-NLVR forward at the bench configuration after warm-up -- the target of the per-kernel ncu captures
(every kernel type of the step appears within the first ViT layer and the first text layer)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from madtp_b200 import synthetic
from madtp_b200.blip_nlvr import BLIP_NLVR, TokenizedText

dev = torch.device("cuda:0")
cal = bench.calibration()
model = BLIP_NLVR(image_size=384, evaluate=True)
model.load_state_dict(synthetic.blip_nlvr_state_dict(1234, img_size=384), strict=False)
model = model.to(dev).eval()
images, ids, mask = synthetic.nlvr_inputs(32, 384, 20, seed=0)
images, text = images.to(dev), TokenizedText(ids.to(dev), mask.to(dev))
# NVTX range "cap" around two ViT blocks and one text layer of the LAST forward:
#   ncu --nvtx --nvtx-include "cap/" ... python scripts/layer_once.py
state = {"on": False}


def _wrap(obj, name):
    fn = getattr(obj, name)

    def wrapped(*a, **k):     # the encoders call these methods directly (not through Module.__call__)
        if not state["on"]:
            return fn(*a, **k)
        torch.cuda.nvtx.range_push("cap")
        try:
            return fn(*a, **k)
        finally:
            torch.cuda.nvtx.range_pop()
    setattr(obj, name, wrapped)


_wrap(model.visual_encoder.blocks[2], "forward_rows")
_wrap(model.visual_encoder.blocks[3], "forward_rows")
_wrap(model.text_encoder.encoder.layer[1], "_forward_impl")
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for it in range(n_iter):
    state["on"] = it == n_iter - 1
    model(images, text, 32, cal["temperature"], train=False)
torch.cuda.synchronize()
