"""Two Flow.train_step calls of CouplingRQNSF(1024, n_hidden=1024) on 16384 rows (BASELINE configs[4], one GPU): the
command profiled by ncu for the wide-conditioner kernels (profiles/r2_wide_*)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchflows_b200 import Flow
from torchflows_b200.architectures import CouplingRQNSF

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device('cuda:0')
torch.manual_seed(0)
flow = Flow(CouplingRQNSF(1024, conditioner_kwargs={'n_hidden': 1024})).to(dev)
x = torch.randn(rows, 1024, device=dev)
flow.train()
flow._optimizer = torch.optim.AdamW(flow.parameters(), lr=1e-3)
for _ in range(steps):
    loss = flow.train_step(x, n_global=rows)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = flow.train_step(x, n_global=rows)
e1.record()
torch.cuda.synchronize()
print('ms per step', e0.elapsed_time(e1) / steps, 'loss', float(loss))
with torch.no_grad():
    flow.eval()
    e0.record()
    lp = flow.log_prob(x)
    e1.record()
    torch.cuda.synchronize()
    print('log_prob ms', e0.elapsed_time(e1), float(lp.mean()))
