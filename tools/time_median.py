import sys, os, torch
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
    for _ in range(10): ops.temporal_median(frames, out=out)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{tag}: {ms:.3f} ms  {(frames.shape[0]+1)*out.numel()/ms/1e6:.0f} GB/s", flush=True)
h, w = 1080, 1920
g = torch.Generator(device=dev).manual_seed(0)
base = torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device=dev, generator=g).to(torch.int16)
for n in (300, 299, 152, 600, 1000, 2000):
    noise = torch.randint(-6, 7, (n, h, w, 3), dtype=torch.int16, device=dev, generator=g)
    frames = (base[None] + noise).clamp_(0, 255).to(torch.uint8); del noise
    timeit(frames, f"noise6 n={n}")
    del frames
frames = bench.make_clip_device(300, h, w, 0, dev)
timeit(frames, "bench clip n=300")
timeit(frames[:299].contiguous(), "bench clip n=299")
rnd = torch.randint(0, 256, (300, h, w, 3), dtype=torch.uint8, device=dev, generator=g)
timeit(rnd, "uniform random n=300")
