import sys; sys.path.insert(0,'/root/repo')
import torch, time
from torchflows_b200 import Flow
from torchflows_b200.architectures import CouplingRQNSF
from torchflows_b200 import _native as N
dev=torch.device('cuda:0')
torch.manual_seed(0)
f = Flow(CouplingRQNSF(256)).to(dev)
x = torch.randn(131072, 256, device=dev)
f.train(); f._optimizer = torch.optim.AdamW(f.parameters(), lr=1e-3)
calls = {'n': 0}
orig = N.wide_coupling_forward
def spy(*a, **k):
    calls['n'] += 1
    return orig(*a, **k)
N.wide_coupling_forward = spy
for _ in range(3): f.train_step(x, n_global=len(x))
torch.cuda.synchronize(); t=time.time()
for _ in range(5): l = f.train_step(x, n_global=len(x))
torch.cuda.synchronize(); print('ms/step', (time.time()-t)/5*1e3, 'wide calls', calls['n'], 'loss', float(l))
for l_ in f.bijection.layers:
    if hasattr(l_, '_wide'): print(l_._fusable, l_._wide_training, l_._trains_wide())
