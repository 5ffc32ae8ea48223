"""Where does a Flow.fit step spend its time?  torch.profiler summary of CouplingRQNSF(256) training steps."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchflows_b200 import Flow  # noqa: E402
from torchflows_b200.architectures import CouplingRQNSF  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
flow = Flow(CouplingRQNSF(256)).to(dev)
x = torch.randn(131072, 256, device=dev)
flow.train()
flow._optimizer = torch.optim.AdamW(flow.parameters(), lr=1e-3)
for _ in range(3):
    flow.train_step(x)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5):
    flow.train_step(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'host time per step {(t1 - t0) / 5 * 1e3:.2f} ms, wall per step {(t2 - t0) / 5 * 1e3:.2f} ms')
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        flow.train_step(x)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=18, max_name_column_width=60))
