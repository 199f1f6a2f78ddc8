"""Host-side floor of one BLIP-NLVR forward: a 2-pair batch makes every kernel tiny, so the wall time per step is the
Python / binding / launch overhead plus the 24 topk read-backs (development aid)."""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cProfile
import pstats
import torch
import bench
from madtp_b200 import synthetic
from madtp_b200.blip_nlvr import BLIP_NLVR, TokenizedText

dev = torch.device("cuda:0")
cal = bench.calibration()
model = BLIP_NLVR(image_size=384, evaluate=True)
model.load_state_dict(synthetic.blip_nlvr_state_dict(1234, img_size=384), strict=False)
model = model.to(dev).eval()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
images, ids, mask = synthetic.nlvr_inputs(P, 384, 20, seed=0)
images, text = images.to(dev), TokenizedText(ids.to(dev), mask.to(dev))
for _ in range(5):
    model(images, text, P, cal["temperature"], train=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 20
for _ in range(n):
    model(images, text, P, cal["temperature"], train=False)
torch.cuda.synchronize()
print(f"pairs={P}: {(time.perf_counter() - t0) / n * 1e3:.2f} ms per step (host-bound)")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    model(images, text, P, cal["temperature"], train=False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
