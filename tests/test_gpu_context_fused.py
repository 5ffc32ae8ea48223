"""Context-conditioned coupling layers on the fused path: the context enters as a per-row hidden bias
(B2F_FLAG_ROW_BIAS; conditioning/context.py:46-60, transforms.py:293-307), and context-conditioned ElementwiseAffine layers carry their predicted
parameters per row, so a whole context-conditioned coupling preset is one flow program again.  Checked against the composite path (conditioner on library GEMMs + stand-alone transformer kernel)
and, in tests/test_gpu_parity.py::test_context_presets_vs_golden, against the reference's own outputs and gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def close(a, b, what, atol, rtol):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.isfinite(a).all(), what
    err = (a - b).abs() - (atol + rtol * b.abs())
    assert (err <= 0).all(), f'{what}: max abs diff {(a - b).abs().max().item():.3e}'


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize('preset,D,ctx_shape,B', [('RealNVP', 8, (3,), 100), ('NICE', 16, (2, 2), 77), ('CouplingRQNSF', 12, (5,), 200),
                                                  ('InverseRealNVP', 7, (3,), 33), ('CouplingRQNSF', 64, (9,), 1000)])
def test_context_couplings_run_fused_and_match_the_composite_path(preset, D, ctx_shape, B):
    from torchflows_b200 import Flow, _native as N
    import torchflows_b200.architectures as arch
    dev = torch.device('cuda:0')
    torch.manual_seed(D)
    flow = Flow(getattr(arch, preset)(D, context_shape=ctx_shape)).to(dev).eval()
    with torch.no_grad():
        for name, p in flow.named_parameters():
            if 'conditioner_transform' in name:
                p.add_(0.3 * torch.randn_like(p))
    x = torch.randn(B, D, device=dev)
    c = torch.randn(B, *ctx_shape, device=dev)
    # the whole context-conditioned preset is ONE program: the couplings carry the context as a per-row hidden bias, the two
    # context-conditioned ElementwiseAffine layers (architectures.py:46,52) as per-row parameters
    segs = flow.bijection._segments('forward', c)
    assert [k for k, _ in segs] == ['ops'], [k for k, _ in segs]
    assert sum(bool(op.flags & N.FLAG_ROW_BIAS) for op in segs[0][1] if op.kind == N.OP_COUPLING) == 2
    assert sum(bool(op.flags & N.FLAG_ROW_BIAS) for op in segs[0][1] if op.kind == N.OP_ELEMENTWISE) == 2

    def run(fused):
        couplings = [l for l in flow.bijection.layers if hasattr(l, '_fusable_ctx')]
        for l in couplings:
            l._fusable_ctx = fused
        for l in flow.bijection.layers:
            if hasattr(l, 'fuse_context'):
                l.fuse_context = fused
        xg, cg = x.clone().requires_grad_(True), c.clone().requires_grad_(True)
        flow.zero_grad()
        z, ld = flow.bijection.forward(xg, context=cg)
        kernel = N.last_flow_kernel()
        lp = flow.log_prob(xg, context=cg)
        xr, ldi = flow.bijection.inverse(z.detach(), context=cg.detach())
        (lp.mean() + 0.1 * z.square().mean()).backward()
        grads = {k: p.grad.clone() for k, p in flow.named_parameters() if p.grad is not None}
        for l in couplings:
            l._fusable_ctx = True
        for l in flow.bijection.layers:
            if hasattr(l, 'fuse_context'):
                l.fuse_context = True
        return z, ld, lp, xr, ldi, xg.grad, cg.grad, grads, kernel

    f = run(True)
    assert f[8] == N.KERNEL_GENERIC          # the whole-flow kernel that takes per-row hidden biases
    g = run(False)
    spline = 'RQNSF' in preset
    # two fp32 conditioners with different summation orders: spline VALUES agree to the spline tolerance (knot noise ~
    # ulp(boundary) amplified by steep bins, cf. tests/test_gpu_rows.py), log-quantities to 1e-4
    close(f[0], g[0], 'z', 2e-3 if spline else 2e-5, 1e-4)
    close(f[1], g[1], 'log_det', 1e-4, 1e-4)
    close(f[2], g[2], 'log_prob', 1e-4, 1e-4)
    close(f[3], x, 'round trip', 2e-3 if spline else 1e-4, 1e-4)
    close(f[4], g[4], 'inverse log_det', 2e-4, 2e-4)
    tol = 5e-3 if spline else 2e-4
    assert rel(f[5], g[5]) < tol, 'd/dx'
    assert rel(f[6], g[6]) < tol, 'd/dcontext'
    for k in g[7]:
        if g[7][k].norm() > 0:
            assert rel(f[7][k], g[7][k]) < tol, k


def test_single_layer_with_context_both_directions():
    from torchflows_b200 import _native as N
    from torchflows_b200.bijections.finite.autoregressive.layers import AffineCoupling
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    layer = AffineCoupling((10,), context_shape=(4,)).to(dev)
    x, c = torch.randn(3, 50, 10, device=dev), torch.randn(3, 50, 4, device=dev)
    with torch.no_grad():
        z, ld = layer.forward(x, context=c)
        assert N.last_flow_kernel() == N.KERNEL_GENERIC
        xr, ldi = layer.inverse(z, context=c)
        zc, ldc = layer._composite(x, c, 'forward')
    assert ld.shape == (3, 50)
    close(z, zc, 'z', 2e-5, 1e-5)
    close(ld, ldc, 'log_det', 1e-5, 1e-5)
    close(xr, x, 'round trip', 1e-4, 1e-4)
    close(ld + ldi, torch.zeros(3, 50), 'antisymmetry', 1e-4, 0)
