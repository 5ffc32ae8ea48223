"""README example (BASELINE configs[0]): RealNVP(3), 1000 standard-normal points, Flow.fit + log_prob + sample(50).
Wall time on the GPU, eager and with fit(cuda_graph=True); any extra argument adds a cProfile of the eager loop."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchflows_b200 import Flow  # noqa: E402
from torchflows_b200.architectures import RealNVP  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
x = torch.randn(1000, 3)
flow = Flow(RealNVP(3)).to(dev)
flow.fit(x[:64], n_epochs=3)          # warm-up: library load, kernels, allocator
torch.cuda.synchronize()
for graph in (False, True):
    torch.manual_seed(1)
    flow = Flow(RealNVP(3)).to(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    flow.fit(x, n_epochs=500, cuda_graph=graph)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    lp = flow.log_prob(x.to(dev))
    s = flow.sample(50)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'cuda_graph={graph}: fit 500 epochs: {t1 - t0:.3f} s ({(t1 - t0) / 500 * 1e3:.3f} ms per epoch), log_prob + sample(50): '
          f'{(t2 - t1) * 1e3:.2f} ms, mean log_prob {lp.mean().item():.4f}')
if len(sys.argv) < 2:
    sys.exit(0)
import cProfile, pstats
flow = Flow(RealNVP(3)).to(dev)
pr = cProfile.Profile()
pr.enable()
flow.fit(x, n_epochs=100)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
