"""Which kernels run, and for how long, for one preset's log_prob / sample?  (torch.profiler kernel table)"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchflows_b200 import Flow  # noqa: E402
from torchflows_b200 import architectures  # noqa: E402

preset, D, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device('cuda:0')
torch.manual_seed(0)
flow = Flow(getattr(architectures, preset)(D)).to(dev)
x = torch.randn(B, D, device=dev)
with torch.no_grad():
    flow.log_prob(x[:4096])          # ActNorm data init
flow.eval()
z = torch.randn(B, D, device=dev)
for _ in range(3):
    with torch.no_grad():
        flow.log_prob(x)
        flow._sample_from_base(z, no_grad=True)
torch.cuda.synchronize()
for name, fn in (('log_prob', lambda: flow.log_prob(x)), ('sample', lambda: flow._sample_from_base(z, no_grad=True))):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        with torch.no_grad():
            for _ in range(5):
                fn()
        torch.cuda.synchronize()
    for e in prof.key_averages():
        if e.device_time_total > 0:
            print(f'{name:9s} {e.key[:70]:70s} n={e.count} avg={e.device_time_total / e.count / 1e3:.3f} ms')
