"""debug helper: remove_invalid_objects on the golden cases, differences against the goldens"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
from make_golden import OBJ_CFGS
from video_unscreen_b200 import ops
from video_unscreen_b200.unscreen.utils.maskprocess import _score_map_dev
g = np.load(os.path.join(ROOT, "tests", "golden", "objects.npz"))
only = [int(v) for v in sys.argv[1:]] or range(6)
for i in only:
    a, seg = g[f"alpha_{i}"], g[f"seg_{i}"]
    h, w = a.shape
    for c, cfg in enumerate(OBJ_CFGS):
        sm = _score_map_dev(h, w, cfg, torch.device("cuda"))
        out, status = ops.remove_invalid_objects(torch.from_numpy(a).cuda(), torch.from_numpy(seg).cuda(), sm,
                                                 cfg['objectremoval']['saliency_thr'], cfg['objectremoval']['consensus_thr'])
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        want = g[f"seg_{i}_{c}"]
        print(i, c, "status", status.cpu().numpy(), "diff px", int((o != want).sum()), "kept want/got", int((want > 0).sum()), int((o > 0).sum()), flush=True)
        if (o != want).any():
            ys, xs = np.nonzero(o != want)
            print("  diff at", list(zip(ys.tolist(), xs.tolist()))[:12])
            y0, x0 = max(ys.min() - 4, 0), max(xs.min() - 4, 0)
            print("  alpha>0 crop at", (y0, x0))
            print((a[y0:y0 + 14, x0:x0 + 24] > 0).astype(int))
            print("  got kept crop"); print((o[y0:y0 + 14, x0:x0 + 24] > 0).astype(int))
            print("  want kept crop"); print((want[y0:y0 + 14, x0:x0 + 24] > 0).astype(int))
            out2, _ = ops.remove_invalid_objects(torch.from_numpy(a).cuda(), torch.from_numpy(seg).cuda(), sm,
                                                 cfg['objectremoval']['saliency_thr'], cfg['objectremoval']['consensus_thr'])
            print("  deterministic:", bool(torch.equal(out, out2)))
