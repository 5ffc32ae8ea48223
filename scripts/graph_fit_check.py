"""fit(cuda_graph=True) against the ordinary eager fit: same data, same initialisation, same batch order."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchflows_b200 import Flow  # noqa: E402
from torchflows_b200.architectures import RealNVP, CouplingRQNSF, MAF  # noqa: E402

dev = torch.device('cuda:0')
for cls, D, n, bs in ((RealNVP, 3, 1000, None), (CouplingRQNSF, 8, 1000, 256), (MAF, 16, 2048, 512), (CouplingRQNSF, 64, 4096, 1024)):
    res = []
    for graph in (False, True):
        torch.manual_seed(0)
        x = torch.randn(n, D) * 1.5 + 0.3
        flow = Flow(cls(D)).to(dev)
        flow.fit(x[:64], n_epochs=2)      # warm the library
        torch.manual_seed(1)
        flow = Flow(cls(D)).to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        flow.fit(x, n_epochs=100, batch_size=bs, lr=0.01, cuda_graph=graph)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        with torch.no_grad():
            res.append((flow.log_prob(x.to(dev)).mean().item(), dt))
    print(f'{cls.__name__}({D}) n={n} batch={bs}: eager {res[0][0]:.5f} in {res[0][1]:.3f} s | graph {res[1][0]:.5f} in {res[1][1]:.3f} s')
