"""Kernel-only durations (torch.profiler) of the bench's alternating log_prob / sample loop on the bench's own flow."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

w = sys.argv[1]
preset, D, B, _, _ = bench.WORKLOADS[w]
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(B, D, device=dev, generator=g)
z = torch.randn(B, D, device=dev, generator=g)
flow = bench.build_flow(preset, D, dev, init_rows=x[:65536])
with torch.no_grad():
    for _ in range(3):
        flow.log_prob(x)
        flow._sample_from_base(z, no_grad=True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            lp = flow.log_prob(x)
            xs = flow._sample_from_base(z, no_grad=True)
        torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_time_total > 0), key=lambda e: e.time_range.start)
for e in evs:
    print(f'{e.name[:60]:60s} start={e.time_range.start / 1e3:10.3f} ms dur={e.device_time_total / 1e3:.3f} ms')
print('xs finite:', bool(torch.isfinite(xs).all()), 'max|xs|', float(xs.abs().max()))
