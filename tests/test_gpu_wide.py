"""Wide-conditioner spline coupling layer (csrc/b2f_wide.cu, include/b2f.h b2f_wide_coupling_*) against the oracle.

The layer is CouplingBijection + FeedForward(n_hidden) + RationalQuadratic of the reference
(layers_base.py:119-163, conditioning/transforms.py:274-307, transformers/spline/rational_quadratic.py:45-200); the
checker is oracle/flow_oracle.py (the reference's ATen ops) in fp32 and fp64 on the CPU.  Tolerances: log_det 1e-4
abs/rel (north star) -- the conditioner GEMMs run in TF32 like the fused whole-flow kernel -- outputs 1e-4 abs/rel.
Gradients: dL/dh of this layer is two orders of magnitude more sensitive to h than the values are (an fp64 evaluation
with only the GEMM operands rounded to TF32 moves the parameter gradients by 1e-2 relative for random upstream gradients,
scripts/wide_diag.py), so the gradient checker is the oracle in fp64 with the *specified* operand rounding (x_A, W1,
tanh(.), W2 to TF32, straight-through), tolerance 2e-3 relative L2 (TF32 operands of the three gradient GEMMs); the
end-to-end check against the unrounded fp64 oracle on the real loss is tests/test_gpu_training.py::
test_gradients_at_benchmark_shapes_with_fp64_referee."""
import pytest
import torch

from oracle import flow_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _params(D, H, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    Dh = D // 2
    W1 = (torch.rand(H, Dh, generator=g) * 2 - 1) / Dh ** 0.5
    b1 = (torch.rand(H, generator=g) * 2 - 1) / Dh ** 0.5
    W2 = (torch.rand(Dh * 23, H, generator=g) * 2 - 1) / H ** 0.5 * scale
    b2 = (torch.rand(Dh * 23, generator=g) * 2 - 1) / H ** 0.5 * scale
    return W1, b1, W2, b2


def _tf32_st(t):
    """Round to TF32 (nearest, ties away: cvt.rna) with a straight-through gradient."""
    r = ((t.detach().float().contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32).to(t.dtype)
    return t + (r - t).detach()


def _oracle_layer(x, W1, b1, W2, b2, direction, boundary, dtype, tf32_operands=False):
    """The reference's coupling layer restated with the oracle's functions: returns (y, log_det).  tf32_operands: the
    arithmetic the kernels specify -- GEMM operands rounded to TF32, everything else in `dtype`."""
    x, W1, b1, W2, b2 = (t.to(dtype) for t in (x, W1, b1, W2, b2))
    Dh = x.shape[1] // 2
    xa, xb = x[:, :Dh], x[:, Dh:]
    rnd = _tf32_st if tf32_operands else (lambda t: t)
    h = (rnd(torch.tanh(rnd(xa) @ rnd(W1).t() + b1)) @ rnd(W2).t() + b2).view(x.shape[0], Dh, 23)
    fn = O.rq_forward if direction == 'forward' else O.rq_inverse
    yb, ld = fn(xb, h, n_bins=8, boundary=boundary)
    return torch.cat([xa, yb], dim=1), ld


@pytest.mark.parametrize('B,D,H,direction', [(300, 64, 32, 'forward'), (300, 64, 32, 'inverse'), (1000, 128, 96, 'forward'),
                                             (513, 192, 256, 'inverse'), (2048, 256, 320, 'forward'), (1, 64, 64, 'forward'),
                                             (700, 256, 17, 'forward'), (300, 64, 5, 'inverse'), (400, 128, 40, 'forward')])
def test_wide_layer_forward_vs_oracle(B, D, H, direction):
    from torchflows_b200 import _native as N
    dev = torch.device('cuda:0')
    W1, b1, W2, b2 = _params(D, H, seed=B + D + H)       # nn.Linear's default init range, like the presets
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, D, generator=g) * 2.0
    x[0, D // 2] = 7.5                       # outside the spline: identity tail
    boundary = 5.0
    tk = N.T_RQ_FWD if direction == 'forward' else N.T_RQ_INV
    y, ld, _ = N.wide_coupling_forward(tk, x.to(dev), *(t.to(dev) for t in (W1, b1, W2, b2)), n_bins=8, boundary=boundary)
    torch.cuda.synchronize()
    y64, ld64 = _oracle_layer(x, W1, b1, W2, b2, direction, boundary, torch.float64)
    assert torch.equal(y[:, :D // 2].cpu(), x[:, :D // 2])
    assert float(y[0, D // 2]) == 7.5
    err_y = (y.cpu().double() - y64).abs() / (1 + y64.abs())
    err_ld = (ld.cpu().double() - ld64).abs() / (1 + ld64.abs())
    assert float(err_y.max()) < 1e-4, float(err_y.max())
    assert float(err_ld.max()) < 1e-4, float(err_ld.max())


@pytest.mark.parametrize('B,D,H,direction', [(300, 64, 32, 'forward'), (700, 128, 96, 'forward'), (515, 128, 64, 'inverse'),
                                             (1024, 256, 288, 'forward'), (600, 256, 17, 'forward'), (300, 64, 9, 'inverse')])
def test_wide_layer_backward_vs_oracle_autograd(B, D, H, direction):
    from torchflows_b200 import _native as N
    dev = torch.device('cuda:0')
    W1, b1, W2, b2 = _params(D, H, seed=3 * B + D + H, scale=2.0)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, D, generator=g) * 2.0
    gy = torch.randn(B, D, generator=g)
    gld = torch.randn(B, generator=g)
    boundary = 5.0
    tk = N.T_RQ_FWD if direction == 'forward' else N.T_RQ_INV
    params = [t.to(dev) for t in (W1, b1, W2, b2)]
    ours = N.wide_coupling_backward(tk, x.to(dev), gy.to(dev), gld.to(dev), *params, n_bins=8, boundary=boundary)
    # same gradients when the backward reuses what the forward of the step kept (packed operands, hidden activations)
    _, _, keep = N.wide_coupling_forward(tk, x.to(dev), *params, n_bins=8, boundary=boundary, for_backward=True)
    kept = N.wide_coupling_backward(tk, x.to(dev), gy.to(dev), gld.to(dev), *params, n_bins=8, boundary=boundary, keep=keep)
    torch.cuda.synchronize()
    for a, b_ in zip(ours, kept):
        assert rel(b_, a) < 1e-4            # same arithmetic; only the order of the split-K and bias-gradient atomics differs

    leaves = [t.double().clone().requires_grad_(True) for t in (x, W1, b1, W2, b2)]
    yo, ldo = _oracle_layer(*leaves, direction, boundary, torch.float64, tf32_operands=True)
    ((yo * gy.double()).sum() + (ldo * gld.double()).sum()).backward()
    for name, o, leaf in zip(('gx', 'gW1', 'gb1', 'gW2', 'gb2'), ours, leaves):
        assert rel(o, leaf.grad) <= 2e-3, f'{name}: ours vs fp64 oracle with TF32 operands {rel(o, leaf.grad):.2e}'


def test_wide_preset_runs_on_the_wide_kernels_and_matches_the_oracle_flow():
    """CouplingRQNSF with a hidden width beyond the whole-flow kernels: every coupling layer reports the wide path, and
    Flow.log_prob / sample agree with OracleFlow on the same state_dict."""
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import CouplingRQNSF
    dev = torch.device('cuda:0')
    torch.manual_seed(5)
    D = 128
    flow = Flow(CouplingRQNSF(D, conditioner_kwargs={'n_hidden': 512})).eval()
    layers = [l for l in flow.bijection.layers if hasattr(l, '_wide')]
    assert layers and all(l._wide and not l._fusable for l in layers)
    sd = {k: v.detach().clone() for k, v in flow.state_dict().items()}
    o = O.OracleFlow('CouplingRQNSF', (D,), {k: v.double() for k, v in sd.items()})
    x = torch.randn(777, D) * 1.5
    flow = flow.to(dev)
    with torch.no_grad():
        lp = flow.log_prob(x.to(dev)).cpu().double()
        lp64 = o.log_prob(x.double())
        assert float(((lp - lp64).abs() / (1 + lp64.abs())).max()) < 1e-4
        z = torch.randn(300, D)
        xs = flow._sample_from_base(z.to(dev), no_grad=True).cpu().double()
        xs64 = o.sample_from_noise(z.double())
        assert rel(xs, xs64) < 1e-4


def test_wide_layer_rejects_unsupported_shapes():
    from torchflows_b200 import _native as N
    dev = torch.device('cuda:0')
    W1, b1, W2, b2 = (t.to(dev) for t in _params(96, 32, seed=0))
    with pytest.raises(N.B2FError):
        N.wide_coupling_forward(N.T_RQ_FWD, torch.zeros(8, 96, device=dev), W1, b1, W2, b2)


# ---- runs of per-column layers (csrc/b2f_colrun.cu) ------------------------------------------------------------------------
def _torch_column_run(kinds, values, x):
    """The reference's layers one by one (affine.py:33-59, permutation.py:19-37) in fp64."""
    import math
    from torchflows_b200 import _native as N
    ld = torch.zeros(x.shape[0], dtype=x.dtype)
    for k, v in zip(kinds, values):
        if k == N.COL_FLIP:
            x = x.flip(1)
            continue
        alpha = torch.exp(math.log(1 - 1e-10) + v[:, 0] / 2) + 1e-10
        if k == N.COL_AFFINE_FWD:
            x = alpha * x + v[:, 1]
            ld = ld + torch.log(alpha).sum()
        else:
            x = (x - v[:, 1]) / alpha
            ld = ld - torch.log(alpha).sum()
    return x, ld


@pytest.mark.parametrize('B,D,kinds', [(100, 64, (0,)), (257, 1024, (0, 2, 1)), (64, 128, (2, 1, 2, 0, 1)), (33, 8, (1, 2, 2, 2, 0, 0, 1, 2)),
                                       (1000, 1024, (1, 0, 1))])
def test_column_run_vs_layer_by_layer_reference(B, D, kinds):
    from torchflows_b200 import _native as N
    from torchflows_b200 import _program as prog
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(B + D)
    values = [None if k == N.COL_FLIP else torch.randn(D, 2, generator=g) for k in kinds]
    x = torch.randn(B, D, generator=g)
    gy, gld = torch.randn(B, D, generator=g), torch.randn(B, generator=g)
    # ours (every second parameter tensor frozen, like ActNorm's)
    vd = [None if v is None else v.to(dev).requires_grad_(i % 2 == 0) for i, v in enumerate(values)]
    xd = x.to(dev).requires_grad_(True)
    y, ld = prog.run_column_ops(list(zip(kinds, vd)), xd)
    ((y * gy.to(dev)).sum() + (ld * gld.to(dev)).sum()).backward()
    # reference
    v64 = [None if v is None else v.double().requires_grad_(True) for v in values]
    x64 = x.double().requires_grad_(True)
    y64, ld64 = _torch_column_run(kinds, v64, x64)
    ((y64 * gy.double()).sum() + (ld64 * gld.double()).sum()).backward()
    assert rel(y, y64) < 1e-6 and float((ld.cpu().double() - ld64).abs().max()) < 1e-4 * (1 + float(ld64.abs().max()))
    assert rel(xd.grad, x64.grad) < 1e-6
    for i, (a, b_) in enumerate(zip(vd, v64)):
        if a is None:
            continue
        if i % 2 == 0:
            assert rel(a.grad, b_.grad) < 1e-4, i
        else:
            assert a.grad is None


def test_wide_flow_edge_batches_and_training_modes():
    """Empty / single-row / ragged batches through a flow whose couplings run on the wide pipeline and whose elementwise
    layers run as column runs (D = 1024: nothing fits the whole-flow kernels), in inference and in a training step; the
    backward that rebuilds its operands (parameters changed between forward and backward) agrees with the kept one."""
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import CouplingRQNSF
    dev = torch.device('cuda:0')
    torch.manual_seed(3)
    flow = Flow(CouplingRQNSF(1024, conditioner_kwargs={'n_hidden': 96})).to(dev)
    assert all(l._wide for l in flow.bijection.layers if hasattr(l, '_wide'))
    flow.eval()
    with torch.no_grad():
        assert flow.log_prob(torch.zeros(0, 1024, device=dev)).shape == (0,)
        x = torch.randn(257, 1024, device=dev)
        lp = flow.log_prob(x)
        assert torch.isfinite(lp).all()
        for n in (1, 255, 256):
            assert torch.allclose(flow.log_prob(x[:n]), lp[:n], rtol=0, atol=1e-3)      # row-independent, any tile raggedness
        xs, lps = flow.sample((3, 5), no_grad=True, return_log_prob=True)
        assert xs.shape == (3, 5, 1024) and lps.shape == (3, 5) and torch.isfinite(xs).all()
        z, ld = flow.bijection.forward(x)
        xr, ldi = flow.bijection.inverse(z)
        assert float((xr - x).abs().max()) < 2e-3 and float((ld + ldi).abs().max()) < 1e-2
    flow.train()
    l0 = float(flow.train_step(x))
    l1 = float(flow.train_step(x))
    assert l0 == l0 and l1 == l1 and l1 < l0 + 1.0
    # gradient w.r.t. the input through wide layers + column runs (variational-style use)
    xg = x.clone().requires_grad_(True)
    flow.log_prob(xg).sum().backward()
    assert torch.isfinite(xg.grad).all() and float(xg.grad.abs().max()) > 0


def test_cluster_and_single_cta_spline_gemms_agree():
    """The 2-CTA cluster kernel (default) and the single-CTA two-tile kernel (B2F_WIDE_NO_CLUSTER=1) run the same arithmetic
    per output: forward results are identical, gradients agree to the order of the atomics."""
    import os
    from torchflows_b200 import _native as N
    dev = torch.device('cuda:0')
    W1, b1, W2, b2 = (t.to(dev) for t in _params(128, 160, seed=4))
    g = torch.Generator().manual_seed(4)
    x = (torch.randn(777, 128, generator=g) * 2).to(dev)
    gy, gld = torch.randn(777, 128, generator=g).to(dev), torch.randn(777, generator=g).to(dev)
    outs = []
    for single in (False, True):
        if single:
            os.environ['B2F_WIDE_NO_CLUSTER'] = '1'
        try:
            y, ld, _ = N.wide_coupling_forward(N.T_RQ_FWD, x, W1, b1, W2, b2, boundary=5.0)
            grads = N.wide_coupling_backward(N.T_RQ_FWD, x, gy, gld, W1, b1, W2, b2, boundary=5.0)
            torch.cuda.synchronize()
            outs.append((y, ld) + tuple(grads))
        finally:
            os.environ.pop('B2F_WIDE_NO_CLUSTER', None)
    assert torch.equal(outs[0][0], outs[1][0])
    assert float((outs[0][1] - outs[1][1]).abs().max()) < 1e-5
    for a, b_ in zip(outs[0][2:], outs[1][2:]):
        assert rel(a, b_) < 1e-4


def test_base_log_prob_kernel_matches_torch_ops():
    """b2f_gauss_log_prob / _backward (base density of flows whose layers are not one program) against
    DiagonalGaussian.log_prob (base_distributions/gaussian.py:46-54) and its autograd."""
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import CouplingRQNSF
    dev = torch.device('cuda:0')
    torch.manual_seed(8)
    flow = Flow(CouplingRQNSF(64)).to(dev)
    with torch.no_grad():
        flow.base.loc.copy_(torch.randn(64, device=dev))
        flow.base.log_scale.copy_(0.3 * torch.randn(64, device=dev))
    for B in (1, 33, 1000):
        z = torch.randn(B, 64, device=dev, requires_grad=True)
        lp = flow.base_log_prob(z)
        g = torch.randn(B, device=dev)
        (lp * g).sum().backward()
        z64 = z.detach().double().requires_grad_(True)
        loc, ls = flow.base.loc.double(), flow.base.log_scale.double()
        ref = (-(0.5 * ((z64 - loc) / ls.exp()) ** 2 + 0.5 * torch.log(torch.tensor(2 * torch.pi, dtype=torch.float64)) + ls)).sum(-1)
        (ref * g.double()).sum().backward()
        assert float((lp.double() - ref).abs().max()) < 1e-4 * (1 + float(ref.abs().max()))
        assert rel(z.grad, z64.grad) < 1e-6


def test_default_width_couplings_can_train_through_the_wide_pipeline():
    """B2F_WIDE_TRAINING=1: spline couplings that fit the whole-flow kernels take the per-layer GEMM pipeline when gradients
    are needed (hidden width padded to the k-block with exact zeros).  Same loss, gradients within the TF32-conditioner
    tolerance of the default (fp32-faithful recompute) path."""
    import os
    from torchflows_b200 import Flow, _native as N
    from torchflows_b200.architectures import CouplingRQNSF
    dev = torch.device('cuda:0')
    torch.manual_seed(12)
    flow = Flow(CouplingRQNSF(128)).to(dev)
    x = torch.randn(2000, 128, device=dev)
    out = []
    for wide in (False, True):
        if wide:
            os.environ['B2F_WIDE_TRAINING'] = '1'
        try:
            flow.zero_grad()
            calls = {'n': 0}
            orig = N.wide_coupling_backward

            def spy(*a, **k):
                calls['n'] += 1
                return orig(*a, **k)
            N.wide_coupling_backward = spy
            try:
                loss = flow._base_batch_loss((x, torch.ones(len(x), device=dev)))
                loss.backward()
            finally:
                N.wide_coupling_backward = orig
            assert (calls['n'] > 0) == wide
            out.append((float(loss.detach()), {k: p.grad.clone() for k, p in flow.named_parameters() if p.grad is not None}))
        finally:
            os.environ.pop('B2F_WIDE_TRAINING', None)
    assert abs(out[0][0] - out[1][0]) <= 1e-4 * (1 + abs(out[0][0]))
    for k, g in out[0][1].items():
        if g.norm() > 0:
            assert rel(out[1][1][k], g) < 2e-2, (k, rel(out[1][1][k], g))
