"""Per-launch device times of the NLVR text encoder alone (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from madtp_b200 import _lib, synthetic
from madtp_b200.blip_nlvr import BLIP_NLVR, TokenizedText

dev = torch.device("cuda:0")
cal = bench.calibration()
model = BLIP_NLVR(image_size=384, evaluate=True)
model.load_state_dict(synthetic.blip_nlvr_state_dict(1234, img_size=384), strict=False)
model = model.to(dev).eval()
images, ids, mask = synthetic.nlvr_inputs(32, 384, 20, seed=0)
images, text = images.to(dev), TokenizedText(ids.to(dev), mask.to(dev))
for _ in range(2):
    model(images, text, 32, cal["temperature"], train=False)
emb = model.last["image_embeds"]
enc = [emb[:32].contiguous(), emb[32:].contiguous()]
idg, mg = ids.to(dev), mask.to(dev)
for _ in range(2):
    model.text_encoder(idg, attention_mask=mg, encoder_hidden_states=enc, space_dict=model.space_dict, temperature=cal["temperature"])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    model.text_encoder(idg, attention_mask=mg, encoder_hidden_states=enc, space_dict=model.space_dict, temperature=cal["temperature"])
e1.record(); torch.cuda.synchronize()
print(f"text encoder: {e0.elapsed_time(e1)/5:.3f} ms per call, image tokens {emb.shape[1]}")
t = _lib.LaunchTimer(); _lib.set_launch_timer(t)
model.text_encoder(idg, attention_mask=mg, encoder_hidden_states=enc, space_dict=model.space_dict, temperature=cal["temperature"])
torch.cuda.synchronize(); _lib.set_launch_timer(None)
tot = 0.0
for i, (name, a, b, meta) in enumerate(t.records):
    ms = a.elapsed_time(b); tot += ms
    if i < 2 or 2 + 31 <= i < 2 + 31 + 31:
        print(f"{i:3d} {name:26s} {str(meta):28s} {ms*1e3:8.1f} us")
print("launches", len(t.records), "sum ms", tot)
