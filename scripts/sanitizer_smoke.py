"""Small end-to-end run of every kernel for compute-sanitizer (memcheck): tiny shapes, odd sizes, tails."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchflows_b200 import Flow  # noqa: E402
import torchflows_b200.architectures as arch  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
for preset, D, B in (('CouplingRQNSF', 32, 130), ('RealNVP', 32, 129), ('NICE', 64, 5), ('MAF', 32, 70),
                     ('MaskedAutoregressiveRQNSF', 32, 33), ('RealNVP', 3, 7), ('CouplingRQNSF', 7, 65), ('IAF', 5, 9)):
    flow = Flow(getattr(arch, preset)(D)).to(dev)
    x = torch.randn(B, D, device=dev)
    flow.train()
    lp = flow.log_prob(x)                    # ActNorm init path + autograd path
    if preset not in ('IAF',):
        lp.mean().backward()
    flow.eval()
    with torch.no_grad():
        flow.log_prob(x)
        flow.sample(B, return_log_prob=True)
        z, _ = flow.bijection.forward(x)
        flow.bijection.inverse(z)
        assert flow.log_prob(x[:0]).shape == (0,)
torch.cuda.synchronize()
print('SANITIZER_SMOKE_DONE')
