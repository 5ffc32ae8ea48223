"""Diagnostic: per-gradient errors of the wide coupling layer against fp64 autograd of the oracle layer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_wide import _params, _oracle_layer, rel
from torchflows_b200 import _native as N

dev = torch.device('cuda:0')
for B, D, H, direction, scale in [(300, 64, 32, 'forward', 1.0), (300, 64, 32, 'forward', 3.0), (700, 128, 96, 'forward', 1.0),
                                  (515, 128, 64, 'inverse', 1.0), (1024, 256, 256, 'forward', 1.0), (1024, 256, 288, 'forward', 1.0)]:
    W1, b1, W2, b2 = _params(D, H, seed=3 * B + D + H, scale=scale)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, D, generator=g) * 2.0
    gy = torch.randn(B, D, generator=g)
    gld = torch.randn(B, generator=g)
    tk = N.T_RQ_FWD if direction == 'forward' else N.T_RQ_INV
    ours = N.wide_coupling_backward(tk, x.to(dev), gy.to(dev), gld.to(dev), *(t.to(dev) for t in (W1, b1, W2, b2)), n_bins=8, boundary=5.0)
    torch.cuda.synchronize()

    def autograd(dtype):
        leaves = [t.to(dtype).clone().requires_grad_(True) for t in (x, W1, b1, W2, b2)]
        y, ld = _oracle_layer(*leaves, direction, 5.0, dtype)
        ((y * gy.to(dtype)).sum() + (ld * gld.to(dtype)).sum()).backward()
        return [t.grad for t in leaves]
    r32, r64 = autograd(torch.float32), autograd(torch.float64)
    leaves = [t.double().clone().requires_grad_(True) for t in (x, W1, b1, W2, b2)]
    yo, ldo = _oracle_layer(*leaves, direction, 5.0, torch.float64, tf32_operands=True)
    ((yo * gy.double()).sum() + (ldo * gld.double()).sum()).backward()
    rt = [t.grad for t in leaves]
    print(B, D, H, direction, scale)
    for name, o, a32, a64, at in zip(('gx', 'gW1', 'gb1', 'gW2', 'gb2'), ours, r32, r64, rt):
        extra = ''
        if name == 'gx':
            extra = f' src half {rel(o[:, :D//2], a64[:, :D//2]):.2e} tgt half {rel(o[:, D//2:], a64[:, D//2:]):.2e}'
        print(f'   {name}: ours vs fp64 {rel(o, a64):.2e}   fp32 vs fp64 {rel(a32, a64):.2e}   ours vs fp64-with-TF32-operands {rel(o, at):.2e}{extra}')
