"""Stand-alone spline transformer kernels (b2f_transformer_apply / _backward) against the HBM roofline: the per-layer figure of
SURVEY 8d.  16384 x 512 elements, dense parameters h (23 floats per element), CUDA-event timing, inputs larger than L2."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torchflows_b200 import _native as N  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
rows, E, P = 16384, 512, 23
x = torch.randn(rows, E, device=dev) * 2
h = torch.randn(rows, E * P, device=dev)
g = torch.randn(rows, E, device=dev)
gld = torch.randn(rows, device=dev)
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms_f = timed(lambda: N.transformer_apply(N.T_RQ_FWD, x, h, E * P, 8, 50.0))
ms_b = timed(lambda: N.transformer_backward(N.T_RQ_FWD, x, h, E * P, g, gld, 8, 50.0))
n = rows * E
for name, ms, bytes_per in (('forward', ms_f, (P + 2) * 4), ('backward', ms_b, (2 * P + 3) * 4)):
    gbs = n * bytes_per / (ms * 1e-3) / 1e9
    print(f'spline transformer {name}: {ms:.3f} ms for {n} elements, {bytes_per} algorithmic B/element -> {gbs:.0f} GB/s = '
          f'{gbs / peak * 100:.1f} % of the HBM peak ({peak:.0f} GB/s)')
