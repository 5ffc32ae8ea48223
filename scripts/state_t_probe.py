"""Probe: error of the training-mode first pass (ActNorm data initialisation) against the reference vectors, per golden case."""
import sys
import torch
sys.path.insert(0, '.')
from tests.test_gpu_parity import build
d = torch.load('tests/golden/presets.pt', weights_only=False)
dev = torch.device('cuda:0')
for i, c in enumerate(d):
    flow = build(c['preset'], c['event_shape'], c['kwargs'], c['state_dict'], dev)
    flow.train()
    with torch.no_grad():
        lp = flow.log_prob(c['x'].to(dev)).cpu().double()
    ref = c['log_prob_T'].double()
    err = ((lp - ref).abs() / (1 + ref.abs())).max().item()
    print(i, c['preset'], tuple(c['x'].shape), f'max |d|/(1+|ref|) = {err:.2e}', f'max abs {(lp - ref).abs().max().item():.2e}', f'|ref| ~ {ref.abs().mean().item():.1f}')
