"""Diagnostic: tca kernel vs generic kernel vs fp64 oracle on one preset."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.flow_oracle import OracleFlow
from torchflows_b200 import Flow, _native as N
import torchflows_b200.architectures as arch

preset, D, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device('cuda:0')
torch.manual_seed(D + 1)
flow = Flow(getattr(arch, preset)(D)).eval()
with torch.no_grad():
    for name, p in flow.named_parameters():
        if name.endswith('.value'):
            p.add_(0.3 * torch.randn_like(p))
o64 = OracleFlow(preset, (D,), {k: v.double() for k, v in flow.state_dict().items()})
flow = flow.to(dev)
g = torch.Generator().manual_seed(B)
x = torch.randn(B, D, generator=g)
with torch.no_grad():
    z, ld = flow.bijection.forward(x.to(dev))
    print('kernel', N.last_flow_kernel())
    os.environ['B2F_DISABLE_ROWS'] = '1'; os.environ['B2F_DISABLE_TC'] = '1'
    zg, ldg = flow.bijection.forward(x.to(dev))
    print('kernel', N.last_flow_kernel())
    z64, ld64 = o64.forward(x.double())
for name, a in (('tca', z), ('generic', zg)):
    e = (a.cpu().double() - z64).abs() / (1 + z64.abs())
    i = int(e.argmax())
    print(name, 'z err vs fp64 max', float(e.max()), 'at', divmod(i, D), 'z64 there', float(z64.reshape(-1)[i]), 'mean', float(e.mean()))
for name, a in (('tca', ld), ('generic', ldg)):
    e = (a.cpu().double() - ld64).abs() / (1 + ld64.abs())
    print(name, 'ld err vs fp64 max', float(e.max()))
