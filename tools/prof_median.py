"""One launch of the temporal median on the bench clip (for ncu): python tools/prof_median.py [n] [h] [w]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from video_unscreen_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
h = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
w = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
dev = torch.device("cuda")
frames = bench.make_clip_device(n, h, w, 0, dev)
out = torch.empty((h, w, 3), dtype=torch.uint8, device=dev)
for _ in range(4):
    ops.temporal_median(frames, out=out)
torch.cuda.synchronize()
print("ok", int(out.sum().item()))
