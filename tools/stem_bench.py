"""Runs the encoder stem (3 scales) a few times for profiling."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rampvo_b200.extractor import MultiScaleMergerDoubleNet
torch.manual_seed(0)
enc = MultiScaleMergerDoubleNet(5, 3).cuda().eval()
ev = torch.randn(1, 1, 5, 480, 640, device="cuda").round()
im = torch.rand(1, 1, 3, 480, 640, device="cuda")
with torch.no_grad(), torch.autocast("cuda", enabled=True):
    for i in range(4):
        enc(events=ev, images=im, mask=torch.tensor([True]), reinit_hidden=(i == 0), out_scale=0.25)
torch.cuda.synchronize()
print("ok")
