"""torch.profiler summary of CouplingRQNSF(1024, n_hidden=1024) training steps (composite path)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchflows_b200 import Flow  # noqa: E402
from torchflows_b200.architectures import CouplingRQNSF  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
flow = Flow(CouplingRQNSF(1024, conditioner_kwargs={'n_hidden': 1024})).to(dev)
x = torch.randn(16384, 1024, device=dev)
flow.train()
flow._optimizer = torch.optim.AdamW(flow.parameters(), lr=1e-3)
for _ in range(3):
    flow.train_step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        flow.train_step(x)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='self_cuda_time_total', row_limit=22, max_name_column_width=70))
