"""Short driver for ncu: a few dense-tracking frames (eggfusion_b200.tracking.DenseTracker.track, 9 Gauss-Newton steps
over a 3-level 1200x680 pyramid).  Usage: python profiles/prof_tracking.py [frames=2]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from eggfusion_b200 import tracking as TRK  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
pm, pf, T0 = bench.tracking_inputs(dev)
trk = TRK.DenseTracker(TRK.TrackingConfig(**bench.TRACK_CFG), dev)
eye = torch.eye(4, device=dev)
for _ in range(frames):
    T, conv = trk.track(pm, pf, T0, eye)
torch.cuda.synchronize()
print("converged", bool(conv), "dense delta t", trk.last_dense_delta[:3, 3].cpu().numpy(), "counts", trk.status.cpu().numpy())
