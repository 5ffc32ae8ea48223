"""Diagnostic: cluster vs single-CTA spline GEMM kernels of the wide layer (forward outputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_wide import _params
from torchflows_b200 import _native as N
dev = torch.device('cuda:0')
for B, D, H in [(300, 64, 32), (300, 64, 64), (1024, 128, 32), (512, 64, 128)]:
    W1, b1, W2, b2 = (t.to(dev) for t in _params(D, H, seed=1))
    x = (torch.randn(B, D, generator=torch.Generator().manual_seed(1)) * 2).to(dev)
    os.environ['B2F_WIDE_NO_CLUSTER'] = '1'
    y0, ld0, _ = N.wide_coupling_forward(N.T_RQ_FWD, x, W1, b1, W2, b2, boundary=5.0)
    torch.cuda.synchronize()
    os.environ.pop('B2F_WIDE_NO_CLUSTER')
    try:
        y1, ld1, _ = N.wide_coupling_forward(N.T_RQ_FWD, x, W1, b1, W2, b2, boundary=5.0)
        torch.cuda.synchronize()
    except Exception as e:
        print(B, D, H, 'cluster kernel failed:', str(e)[:200]); break
    dy = (y1 - y0).abs()
    print('  per-column mean |dy| of row block 128..255:', [round(float(v), 6) for v in dy[128:256, D // 2:].mean(dim=0)[:16]], ' rows 128..135 max', [round(float(v), 6) for v in dy[128:136].max(dim=1).values])
    bad_rows = (dy.max(dim=1).values > 1e-5).nonzero().flatten()
    bad_cols = (dy.max(dim=0).values > 1e-5).nonzero().flatten()
    print(B, D, H, 'max dy', float(dy.max()), 'max dld', float((ld1 - ld0).abs().max()), 'bad rows', bad_rows[:8].tolist(), len(bad_rows), 'bad cols', bad_cols[:12].tolist(), len(bad_cols))
