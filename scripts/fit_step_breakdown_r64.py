"""Flow.fit step of RealNVP(64) / MAF(128), 131072 rows: kernel table."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchflows_b200 import Flow  # noqa: E402
from torchflows_b200 import architectures  # noqa: E402

dev = torch.device('cuda:0')
for preset, D in (('RealNVP', 64), ('MAF', 128)):
    torch.manual_seed(0)
    flow = Flow(getattr(architectures, preset)(D)).to(dev)
    x = torch.randn(131072, D, device=dev)
    flow.train()
    flow._optimizer = torch.optim.AdamW(flow.parameters(), lr=1e-3)
    for _ in range(3):
        flow.train_step(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            flow.train_step(x)
        torch.cuda.synchronize()
    print(preset, D)
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:5]:
        print(f'   {e.key[:70]:70s} n={e.count} avg={e.device_time_total / e.count / 1e3:.3f} ms')
