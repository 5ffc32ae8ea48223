"""Timing probe: sequential-direction sampling of MaskedAutoregressiveRQNSF(128) / MAF(128), 2^18 rows, per math mode."""
import sys
import torch
sys.path.insert(0, '.')
from torchflows_b200 import Flow, _program as P
import torchflows_b200.architectures as arch

dev = torch.device('cuda:0')
for preset in ('MaskedAutoregressiveRQNSF', 'MAF'):
    torch.manual_seed(0)
    flow = Flow(getattr(arch, preset)(128)).to(dev).eval()
    z = torch.randn(1 << 18, 128, device=dev)
    for mode in ('default', 'fast'):
        prev = P.set_math_mode(mode)
        with torch.no_grad():
            for _ in range(3):
                flow._sample_from_base(z, no_grad=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                flow._sample_from_base(z, no_grad=True)
            e1.record()
            torch.cuda.synchronize()
        P.set_math_mode(prev)
        print(preset, mode, f'{e0.elapsed_time(e1) / 10:.3f} ms')
