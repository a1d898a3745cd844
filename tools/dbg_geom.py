import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_unscreen_b200 import ops
from oracle import cvmodel as M
what, shape, dx, dy = sys.argv[1], eval(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4])
rng = np.random.default_rng(0)
clip = rng.integers(0, 256, (int(os.environ.get("DBG_N", "2")),) + shape, dtype=np.uint8)
ch = 3 if len(shape) == 3 else 1
if what == "shift":
    got = ops.shift(torch.from_numpy(clip).cuda(), dx, dy, ch); torch.cuda.synchronize()
    want = np.stack([M.warp_translate(f, dx, dy) for f in clip])
else:
    got = ops.rescale_cubic(torch.from_numpy(clip).cuda(), dx, ch); torch.cuda.synchronize()
    want = np.stack([M.resize_cubic_crop(f, dx) for f in clip])
got = got.cpu().numpy()
print(what, shape, dx, dy, "mismatch", int((got != want).sum()), "of", got.size, flush=True)
