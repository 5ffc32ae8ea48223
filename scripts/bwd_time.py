"""CUDA-event time of loss.backward() alone for CouplingRQNSF(256), 131072 rows (device idle before, sync after)."""
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
for it in range(6):
    flow.zero_grad(set_to_none=True)
    loss = -flow.log_prob(x).mean()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss.backward()
    e1.record()
    torch.cuda.synchronize()
    print(f'iter {it}: backward {e0.elapsed_time(e1):.3f} ms')
