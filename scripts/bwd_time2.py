"""backward time with / without parameter gradients (atomics) -- CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchflows_b200 import Flow  # noqa: E402
from torchflows_b200.architectures import CouplingRQNSF  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
flow = Flow(CouplingRQNSF(256)).to(dev)
x = torch.randn(131072, 256, device=dev)
flow.train()
with torch.no_grad():
    flow.log_prob(x)
for mode in ('params', 'input-only', 'params'):
    for p in flow.parameters():
        p.requires_grad_(mode == 'params')
    xx = x.clone().requires_grad_(mode != 'params')
    for it in range(4):
        flow.zero_grad(set_to_none=True)
        loss = -flow.log_prob(xx).mean()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss.backward()
        e1.record()
        torch.cuda.synchronize()
    print(f'{mode}: backward {e0.elapsed_time(e1):.3f} ms')
