import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_unscreen_b200 import ops
import bench
dev = torch.device("cuda")
def timeit(frames, tag):
    out = torch.empty(frames.shape[1:], dtype=torch.uint8, device=dev)
    for _ in range(3): ops.temporal_median(frames, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.temporal_median(frames, out=out)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{tag}: {ms:.3f} ms  {(frames.shape[0]+1)*out.numel()/ms/1e6:.0f} GB/s", flush=True)
    return out
frames = bench.make_clip_device(300, 1080, 1920, 0, dev)
a = timeit(frames, "bench clip n=300")
b = timeit(frames[:299].contiguous(), "bench clip n=299")
s, _ = frames[:, 500:504].to(torch.int16).sort(0)
assert torch.equal(a[500:504], ((s[149] + s[150]) >> 1).to(torch.uint8))
print("ok")
