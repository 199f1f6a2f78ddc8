"""Host-side profile of one BLIP-NLVR forward (development aid): where does the Python/ctypes launch time go?"""
import cProfile
import pstats
import sys
import time
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
for _ in range(3):
    model(images, text, 32, cal["temperature"], train=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    model(images, text, 32, cal["temperature"], train=False)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue time per step {(t1 - t0) / 5 * 1e3:.2f} ms; wall per step {(t2 - t0) / 5 * 1e3:.2f} ms")
# text encoder alone
emb = model.last["image_embeds"]
enc = [emb[:32].contiguous(), emb[32:].contiguous()]
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    model.text_encoder(ids.to(dev), attention_mask=mask.to(dev), encoder_hidden_states=enc, space_dict=model.space_dict,
                       temperature=cal["temperature"])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"text encoder: host enqueue {(t1 - t0) / 5 * 1e3:.2f} ms; wall {(t2 - t0) / 5 * 1e3:.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    model(images, text, 32, cal["temperature"], train=False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
