"""Host-side cost of one Flow.log_prob / sample call (tiny batch: the GPU work is negligible)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchflows_b200 import Flow  # noqa: E402
from torchflows_b200 import architectures  # noqa: E402

preset, D = sys.argv[1], int(sys.argv[2])
dev = torch.device('cuda:0')
flow = Flow(getattr(architectures, preset)(D)).to(dev).eval()
x = torch.randn(64, D, device=dev)
with torch.no_grad():
    for _ in range(20):
        flow.log_prob(x)
        flow._sample_from_base(x, no_grad=True)
    torch.cuda.synchronize()
    for name, fn in (('log_prob', lambda: flow.log_prob(x)), ('sample', lambda: flow._sample_from_base(x, no_grad=True))):
        t0 = time.perf_counter()
        for _ in range(2000):
            fn()
        torch.cuda.synchronize()
        print(f'{name}: {(time.perf_counter() - t0) / 2000 * 1e6:.1f} us per call')
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(500):
        flow.log_prob(x)
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
