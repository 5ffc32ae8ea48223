"""Randomised parity sweep on the GPU (not part of the test suite): random presets / event sizes / batch sizes / depths /
weight states against the CPU oracle, and context-conditioned presets fused against composite.  Prints every violation.

    python scripts/fuzz_parity.py [seconds] [seed] [grad|big]

With `grad`: the training loss (flows.py:199-224) and all its gradients against torch autograd through the fp64 oracle
(context-conditioned presets: fused against composite), relative L2 per tensor.
"""
import sys
import time

import torch

sys.path.insert(0, '.')
from oracle.flow_oracle import OracleFlow  # noqa: E402
from torchflows_b200 import Flow  # noqa: E402
import torchflows_b200.architectures as arch  # noqa: E402

PRESETS = ['NICE', 'RealNVP', 'InverseRealNVP', 'MAF', 'IAF', 'CouplingRQNSF', 'MaskedAutoregressiveRQNSF',
           'InverseAutoregressiveRQNSF', 'CouplingLRS', 'MaskedAutoregressiveLRS', 'InverseAutoregressiveLRS']
dev = torch.device('cuda:0')


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).abs() / (1 + b.abs())).max().item() if a.numel() else 0.0


def rel_l2(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def grad_case(preset, D, B, n_layers, ctx_shape):
    """max relative-L2 gradient error over dloss/dx and every parameter with a non-negligible reference gradient."""
    kw = dict(context_shape=ctx_shape) if ctx_shape else {}
    flow = Flow(getattr(arch, preset)(D, n_layers=n_layers, **kw)).to(dev).eval()
    with torch.no_grad():
        for name, p in flow.named_parameters():
            if 'conditioner_transform' in name:
                p.add_(0.2 * torch.randn_like(p))
    x = torch.randn(B, D) * 1.2
    c = torch.randn(B, *ctx_shape).to(dev) if ctx_shape else None

    def ours(fused=True):
        for l in flow.bijection.layers:
            if hasattr(l, '_fusable_ctx'):
                l._fusable_ctx = fused
            if hasattr(l, 'fuse_context'):
                l.fuse_context = fused
        flow.zero_grad()
        xg = x.to(dev).requires_grad_(True)
        batch = (xg, torch.ones(B, device=dev)) + ((c,) if ctx_shape else ())
        loss = flow._base_batch_loss(batch)
        loss.backward()
        return float(loss), xg.grad.clone(), {k: p.grad.clone() for k, p in flow.named_parameters() if p.grad is not None}
    loss, gx, grads = ours(True)
    if ctx_shape:
        loss_ref, gx_ref, grads_ref = ours(False)
    else:
        leaves = {k: v.detach().cpu().double().requires_grad_(True) if v.is_floating_point() and v.numel() > 0 else v.cpu().clone()
                  for k, v in flow.state_dict().items()}
        o = OracleFlow(preset, (D,), {}, n_layers=n_layers)
        o.sd = leaves
        xr = x.double().requires_grad_(True)
        lref = o.batch_loss(xr)
        lref.backward()
        loss_ref, gx_ref = float(lref), xr.grad
        grads_ref = {k: v.grad for k, v in leaves.items() if isinstance(v, torch.Tensor) and v.requires_grad and v.grad is not None}
    worst = rel_l2(gx, gx_ref)
    scale = max(float(g.norm()) for g in grads_ref.values())
    for k, g in grads_ref.items():
        if k in grads and float(g.norm()) > 1e-3 * scale:
            worst = max(worst, rel_l2(grads[k], g))
    return abs(loss - loss_ref) / (1 + abs(loss_ref)), worst


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    if len(sys.argv) > 3 and sys.argv[3] == 'grad':
        return main_grad(budget, int(sys.argv[2]))
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    g = torch.Generator().manual_seed(seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    big = len(sys.argv) > 3 and sys.argv[3] == 'big'
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < budget:
        preset = PRESETS[ri(0, len(PRESETS) - 1)]
        spline = 'RQNSF' in preset or 'LRS' in preset
        seq = preset.startswith(('MAF', 'IAF', 'Masked', 'InverseAuto'))
        D = [2, 3, 5, 8, 16, 24, 32, 33, 48, 64, 96, 128, 160][ri(0, 12)]
        if 'LRS' in preset and seq:
            D = min(D, 16)              # composite D-step loop in Python
        B = [1, 7, 100, 129, 1000, 3000, 20000][ri(0, 6 if D <= 64 and not seq else 4)]
        if big and D % 32 == 0 and ri(0, 2) == 0:
            B = [19077, 40000 + ri(0, 127)][ri(0, 1)]          # more 128-row tiles than SMs, ragged tail (persistent kernels)
        n_layers = ri(1, 3)
        state = 'ET'[ri(0, 1)]
        ctx = ri(0, 3) == 0 and 'LRS' not in preset
        torch.manual_seed(ri(0, 10 ** 6))
        kw = dict(context_shape=(ri(1, 5),)) if ctx else {}
        try:
            flow = Flow(getattr(arch, preset)(D, n_layers=n_layers, **kw)).to(dev)
            x = torch.randn(B, D) * 1.3
            z = torch.randn(B, D)
            c = torch.randn(B, *kw['context_shape']).to(dev) if ctx else None
            if state == 'T' and B > 1:
                flow.train()
                with torch.no_grad():
                    flow.log_prob(x.to(dev), context=c) if ctx else flow.log_prob(x.to(dev))
            flow.eval()
            with torch.no_grad():
                if ctx:
                    lp = flow.log_prob(x.to(dev), context=c)
                    xs, ld = flow.bijection.inverse(z.to(dev), context=c)
                    for l in flow.bijection.layers:
                        if hasattr(l, '_fusable_ctx'):
                            l._fusable_ctx = False
                        if hasattr(l, 'fuse_context'):
                            l.fuse_context = False
                    lp_ref = flow.log_prob(x.to(dev), context=c)
                    xs_ref, ld_ref = flow.bijection.inverse(z.to(dev), context=c)
                    e_lp, e_xs, e_ld = rel(lp, lp_ref), rel(xs, xs_ref), rel(ld, ld_ref)
                else:
                    lp = flow.log_prob(x.to(dev))
                    xs, lps = flow._sample_from_base(z.to(dev), no_grad=True, return_log_prob=True)
                    o = OracleFlow(preset, (D,), {k: v.cpu() for k, v in flow.state_dict().items()}, n_layers=n_layers)
                    chunks = range(0, B, 4096)
                    lp_ref = torch.cat([o.log_prob(x[i:i + 4096]) for i in chunks])
                    outs = [o.sample_from_noise(z[i:i + 4096], return_log_prob=True) for i in chunks]
                    xs_ref, lps_ref = torch.cat([a for a, _ in outs]), torch.cat([b for _, b in outs])
                    e_lp, e_xs, e_ld = rel(lp, lp_ref), rel(xs, xs_ref), rel(lps, lps_ref)
            tol_x = 3e-3 if spline else 1e-4
            ok = e_lp < 1e-4 and e_xs < tol_x and e_ld < 2e-4 and bool(torch.isfinite(lp).all())
        except Exception as e:      # noqa: BLE001
            ok, e_lp, e_xs, e_ld = False, -1, -1, -1
            print('EXCEPTION', preset, D, B, n_layers, state, ctx, repr(e)[:200], flush=True)
        n += 1
        if n <= 12:
            print(f'  e.g. {preset} D={D} B={B} layers={n_layers} state={state} ctx={ctx}: log_prob {e_lp:.2e} sample {e_xs:.2e} {e_ld:.2e}', flush=True)
        if not ok:
            bad += 1
            print(f'VIOLATION {preset} D={D} B={B} layers={n_layers} state={state} ctx={ctx}: log_prob {e_lp:.2e} sample {e_xs:.2e} '
                  f'sample-lp/ld {e_ld:.2e}', flush=True)
    print(f'{n} random configurations, {bad} violations, {time.time() - t0:.0f} s')


def main_grad(budget, seed):
    g = torch.Generator().manual_seed(seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    presets = [p for p in PRESETS if p not in ('InverseAutoregressiveRQNSF',)]      # sequential spline density: exact-log-det flag
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < budget:
        preset = presets[ri(0, len(presets) - 1)]
        spline = 'RQNSF' in preset or 'LRS' in preset
        D = [2, 3, 5, 8, 16, 24, 32, 33, 64][ri(0, 8)]
        if 'LRS' in preset and preset != 'CouplingLRS':
            D = min(D, 8)
        B = [7, 64, 300, 1000][ri(0, 3)]
        ctx_shape = (ri(1, 4),) if (ri(0, 2) == 0 and 'LRS' not in preset) else None
        torch.manual_seed(ri(0, 10 ** 6))
        try:
            e_loss, e_g = grad_case(preset, D, B, ri(1, 2), ctx_shape)
            ok = e_loss < 1e-4 and e_g < (2e-2 if spline else 1e-3)
        except Exception as e:      # noqa: BLE001
            ok, e_loss, e_g = False, -1, -1
            print('EXCEPTION', preset, D, B, ctx_shape, repr(e)[:200], flush=True)
        n += 1
        if n <= 10:
            print(f'  e.g. {preset} D={D} B={B} ctx={ctx_shape}: loss {e_loss:.2e} worst gradient {e_g:.2e}', flush=True)
        if not ok:
            bad += 1
            print(f'VIOLATION {preset} D={D} B={B} ctx={ctx_shape}: loss {e_loss:.2e} worst gradient {e_g:.2e}', flush=True)
    print(f'{n} random gradient configurations, {bad} violations, {time.time() - t0:.0f} s')


if __name__ == '__main__':
    main()
