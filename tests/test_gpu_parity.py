"""GPU parity tests proper: the CUDA path (through the C ABI, via the torchflows_b200 classes) against the oracle
and against the reference-generated golden vectors.  Run on the B200 box: pytest -m gpu."""
import math

import numpy as np
import pytest
import torch

from oracle import c_oracle, flow_oracle as fo

pytestmark = pytest.mark.gpu

LP_TOL = 1e-4      # north_star: per-sample log_prob within 1e-4 abs/rel in fp32


def z_atol(boundary):
    """see tests/test_c_oracle_and_hostmath.py: fp32 knot rounding is worth a few ulp(boundary)."""
    return 1e-5 + 16 * float(np.spacing(np.float32(boundary)))


def close(a, b, what, atol, rtol):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs() - (atol + rtol * b.abs())
    assert torch.isfinite(a).all(), what
    assert (err <= 0).all(), f'{what}: max abs diff {(a - b).abs().max().item():.3e}, worst excess {err.max().item():.3e}'


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'these tests need the B200'
    return torch.device('cuda:0')


def build(preset, event_shape, kwargs, state_dict, dev):
    import torchflows_b200.architectures as arch
    from torchflows_b200 import Flow
    flow = Flow(getattr(arch, preset)(event_shape, **kwargs))
    flow.load_state_dict(state_dict)
    return flow.to(dev).eval()


# ---------------------------------------------------------------------------------------------------------
# P1: transformer level (identical h on both sides)
# ---------------------------------------------------------------------------------------------------------
def test_transformers_vs_golden_and_oracle(golden, dev):
    from torchflows_b200.bijections.finite.autoregressive.transformers.linear.affine import Affine, InverseAffine, Shift
    from torchflows_b200.bijections.finite.autoregressive.transformers.spline.rational_quadratic import RationalQuadratic
    n_rq = 0
    for c in golden('transformers.pt'):
        x, h = c['x'].to(dev), c['h'].to(dev)
        if c['kind'] == 'rq':
            tr = RationalQuadratic(c['event_shape'], boundary=c['boundary'], n_bins=c['n_bins'])
            za = z_atol(c['boundary'])
        else:
            tr = {'affine': Affine, 'inverse_affine': InverseAffine, 'shift': Shift}[c['kind']](c['event_shape'])
            za = 1e-5
        z, ld = tr.forward(x, h)
        xr, ldi = tr.inverse(c['z'].to(dev), h)
        close(z, c['z'], f"{c['kind']} z", za, 1e-5)
        close(ld, c['ld_f'], 'ld_f', LP_TOL, LP_TOL)
        close(xr, c['xr'], 'xr', za, 1e-5)
        close(ldi, c['ld_i'], 'ld_i', LP_TOL, LP_TOL)
        assert z.shape == x.shape and ld.shape == c['ld_f'].shape
        if c['kind'] == 'rq':
            n_rq += 1
            # bin indices: bit-exact against the C oracle (same deterministic knot arithmetic) ...
            kf = tr.bin_indices(x, h, inverse=False).cpu()
            ki = tr.bin_indices(c['z'].to(dev), h, inverse=True).cpu()
            _, _, kf_o = c_oracle.rq(c['x'], c['h'], c['n_bins'], c['boundary'], False)
            _, _, ki_o = c_oracle.rq(c['z'], c['h'], c['n_bins'], c['boundary'], True)
            assert torch.equal(kf, kf_o) and torch.equal(ki, ki_o)
            # ... and equal to the reference's own searchsorted result on these vectors
            assert torch.equal(kf.long(), c['k_f']) and torch.equal(ki.long(), c['k_i'])
            # out-of-bounds elements come back bit-identical with zero log-det contribution
            oob = ~((c['x'] > -c['boundary']) & (c['x'] < c['boundary']))
            assert torch.equal(z.cpu()[oob], c['x'][oob])
    assert n_rq > 40


def test_rq_bins_bit_exact_large(dev):
    """2^16 x 128 elements, h ~ randn: bin indices identical to the C oracle, values within tolerance."""
    from torchflows_b200.bijections.finite.autoregressive.transformers.spline.rational_quadratic import RationalQuadratic
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1 << 13, 128, generator=g) * 4
    h = torch.randn(1 << 13, 128, 23, generator=g)
    tr = RationalQuadratic((128,))
    for inverse in (False, True):
        k = tr.bin_indices(x.to(dev), h.to(dev), inverse=inverse).cpu()
        out, ld = (tr.inverse if inverse else tr.forward)(x.to(dev), h.to(dev))
        o_o, ld_o, k_o = c_oracle.rq(x, h, 8, 50.0, inverse)
        assert torch.equal(k, k_o)
        close(out, o_o, 'out', z_atol(50.0), 1e-5)
        close(ld, ld_o.sum(-1), 'ld', LP_TOL, LP_TOL)


def test_rq_identity_at_zero_parameters(dev):
    from torchflows_b200.bijections.finite.autoregressive.transformers.spline.rational_quadratic import RationalQuadratic
    tr = RationalQuadratic((5,))
    x = torch.randn(11, 5, device=dev)
    z, ld = tr.forward(x, torch.zeros(11, 5, 23, device=dev))
    assert torch.allclose(z, x, atol=1e-2) and torch.allclose(ld, torch.zeros_like(ld), atol=1e-2)


# ---------------------------------------------------------------------------------------------------------
# P2 / P3: preset level, fused whole-flow kernel
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('idx', range(16))
def test_presets_vs_golden(golden, dev, idx):
    c = golden('presets.pt')[idx]
    flow = build(c['preset'], c['event_shape'], c['kwargs'], c['state_dict'], dev)
    rq = 'RQNSF' in c['preset']
    za = z_atol(50.0) * 4 if rq else 2e-5       # two spline layers + elementwise scales in between
    with torch.no_grad():
        z, ld = flow.bijection.forward(c['x'].to(dev))
        lp = flow.log_prob(c['x'].to(dev))
        z2, lp2 = flow.forward_with_log_prob(c['x'].to(dev))
        xs, lps = flow._sample_from_base(c['noise'].to(dev), no_grad=True, return_log_prob=True)
        xr, ldr = flow.bijection.inverse(z)
    assert z.shape == c['x'].shape and ld.shape == c['ld_f'].shape and lp.shape == c['log_prob'].shape
    close(z, c['z'], 'z', za, 1e-4)
    close(ld, c['ld_f'], 'ld_f', LP_TOL, LP_TOL)
    close(lp, c['log_prob'], 'log_prob', LP_TOL, LP_TOL)
    close(lp2, c['log_prob'], 'log_prob (forward_with_log_prob)', LP_TOL, LP_TOL)
    close(z2, c['z'], 'z2', za, 1e-4)
    close(xs, c['xs'], 'sample x', za * 4, 1e-4)
    close(lps, c['lp_s'], 'sample log_prob', 2 * LP_TOL, 2 * LP_TOL)
    # round trip with our own forward output (P3): reference floor is 1e-3 (test/constants.py), its own fp32
    # error is 2e-6 (affine) ... 1.8e-4 (spline)
    rt = (xr.cpu() - c['x']).abs().max().item()
    rt_ref = (c['xr'] - c['x']).abs().max().item()
    assert rt <= max(1e-5, 2.0 * rt_ref), (rt, rt_ref)
    assert (ldr.cpu() + ld.cpu()).abs().max().item() <= max(1e-4, 2.0 * (c['ld_r'] + c['ld_f']).abs().max().item())


@pytest.mark.parametrize('idx', [0, 5, 6, 9, 11, 12])
def test_actnorm_data_init_state_T(golden, dev, idx):
    """First training-mode forward initialises every ActNorm from the activations at its depth (layers.py:58-68)."""
    c = golden('presets.pt')[idx]
    flow = build(c['preset'], c['event_shape'], c['kwargs'], c['state_dict'], dev)
    flow.train()
    with torch.no_grad():
        lp = flow.log_prob(c['x'].to(dev))
    sd = flow.state_dict()
    for k, v in c['actnorm_T'].items():
        close(sd[k], v, k, 2e-4, 2e-4)
    close(lp, c['log_prob_T'], 'log_prob_T', LP_TOL, LP_TOL)        # measured: <= 3.4e-6 relative on all 16 cases (scripts/state_t_probe.py)


def test_oracle_agrees_on_fresh_inputs(golden, dev):
    """Same seeded inputs through the CUDA path and the CPU oracle (not only the stored vectors)."""
    for idx in (9, 11, 12, 14):
        c = golden('presets.pt')[idx]
        flow = build(c['preset'], c['event_shape'], c['kwargs'], c['state_dict'], dev)
        o = fo.OracleFlow(c['preset'], c['event_shape'], c['state_dict'])
        g = torch.Generator().manual_seed(17 + idx)
        x = torch.randn(300, *c['event_shape'], generator=g) * 1.5
        with torch.no_grad():
            lp = flow.log_prob(x.to(dev))
        close(lp, o.log_prob(x), f"{c['preset']} log_prob", LP_TOL, LP_TOL)


def test_shapes_inputs_not_mutated_and_cpu_input(golden, dev):
    c = golden('presets.pt')[2]     # RealNVP (3,5,2), batch (5,)
    flow = build(c['preset'], c['event_shape'], c['kwargs'], c['state_dict'], dev)
    x = c['x'].to(dev)
    x0 = x.clone()
    with torch.no_grad():
        z, ld = flow.bijection.forward(x)
        lp_cpu_input = flow.log_prob(c['x'])           # CPU data is moved like flows.py:646
    assert torch.equal(x, x0) and z.data_ptr() != x.data_ptr()
    close(lp_cpu_input, c['log_prob'], 'lp', LP_TOL, LP_TOL)
    with pytest.raises(Exception):
        flow.bijection.forward(c['x'])                 # a CPU tensor on the kernel path is an error, not a fallback


def test_full_size_properties(dev):
    """BASELINE sizes through size-independent properties: round trip, log-det antisymmetry, sample log-prob
    identity log_prob(x) == base(z) - ld_inv, batch-slicing invariance."""
    import torchflows_b200.architectures as arch
    from torchflows_b200 import Flow
    torch.manual_seed(0)
    for preset, D, B, rt_tol in (('RealNVP', 64, 1 << 20, 1e-4), ('CouplingRQNSF', 256, 1 << 18, 2e-3),
                                 ('MAF', 128, 1 << 16, 1e-4), ('MaskedAutoregressiveRQNSF', 128, 1 << 14, 2e-3)):
        flow = Flow(getattr(arch, preset)(D)).to(dev).eval()
        x = torch.randn(B, D, device=dev)
        with torch.no_grad():
            z, ld = flow.bijection.forward(x)
            xr, ldi = flow.bijection.inverse(z)
            lp = flow.log_prob(x)
            lp_slice = flow.log_prob(x[1000:1777])
        assert torch.isfinite(z).all() and torch.isfinite(lp).all()
        assert (xr - x).abs().max().item() < rt_tol, preset
        if 'MaskedAutoregressiveRQNSF' != preset:     # the reference's sequential log-det quirk breaks antisymmetry
            assert (ld + ldi).abs().max().item() < 5e-3 * max(1.0, ld.abs().max().item() * 1e-2), preset
        assert torch.equal(lp[1000:1777], lp_slice)
        base = -0.5 * (z.double() ** 2).sum(-1) - 0.5 * D * math.log(2 * math.pi)
        assert torch.allclose(lp.double(), base + ld.double(), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize('idx', range(7))
def test_context_presets_vs_golden(golden, dev, idx):
    """Context-conditioned presets against the reference: values, log-densities, round trip, gradients."""
    import torchflows_b200.architectures as arch
    from torchflows_b200 import Flow
    c = golden('context.pt')[idx]
    flow = Flow(getattr(arch, c['preset'])(c['event_shape'], context_shape=c['context_shape']))
    flow.load_state_dict(c['state_dict'])
    flow = flow.to(dev).eval()
    x, ctx = c['x'].to(dev), c['context'].to(dev)
    rq = 'RQNSF' in c['preset']
    za = z_atol(50.0) * 4 if rq else 2e-5
    with torch.no_grad():
        z, ld = flow.bijection.forward(x, context=ctx)
        lp = flow.log_prob(x, context=ctx)
        xr, ldr = flow.bijection.inverse(c['z'].to(dev), context=ctx)
    close(z, c['z'], 'z', za, 1e-4)
    close(ld, c['ld_f'], 'ld_f', LP_TOL, LP_TOL)
    close(lp, c['log_prob'], 'log_prob', LP_TOL, LP_TOL)
    close(xr, c['xr'], 'xr', za * 4, 1e-4)
    close(ldr, c['ld_r'], 'ld_r', 2 * LP_TOL, 2 * LP_TOL)
    if c['preset'] == 'InverseAutoregressiveRQNSF':
        return           # spline + sequential density direction: fused gradient only for the exact log-det flag
    ne = len(c['event_shape'])
    xg = x.reshape(-1, *c['event_shape']).clone().requires_grad_(True)
    loss = flow._base_batch_loss((xg, torch.ones(len(xg), device=dev), ctx.reshape(-1, *c['context_shape'])))
    loss.backward()
    assert abs(float(loss.detach()) - float(c['loss'])) <= 1e-5 * (1 + abs(float(c['loss'])))
    tol = 5e-3 if rq else 1e-4      # tiny batches: the fp32 spline-knot gradient noise (6e-4 in the reference) is not averaged

    def rel(a, b):
        a, b = a.detach().cpu().double(), b.double()
        return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    assert rel(xg.grad, c['grad_x']) < tol
    params = dict(flow.named_parameters())
    for k, g in c['grads'].items():
        if g.norm() > 0:
            assert rel(params[k].grad, g) < tol, k
        elif params[k].grad is not None:
            assert params[k].grad.abs().max().item() < 1e-6, k


def test_empty_batch_and_odd_shapes(dev):
    """Edge cases: empty batch, single row, odd D, batch not a multiple of the tile, through every kernel family."""
    import torchflows_b200.architectures as arch
    from torchflows_b200 import Flow
    torch.manual_seed(0)
    for preset, D, B in (('CouplingRQNSF', 32, 130), ('RealNVP', 32, 129), ('NICE', 64, 5), ('MAF', 32, 70),
                         ('MaskedAutoregressiveRQNSF', 32, 33), ('RealNVP', 3, 7), ('CouplingRQNSF', 7, 65), ('IAF', 5, 1)):
        flow = Flow(getattr(arch, preset)(D)).to(dev).eval()
        x = torch.randn(B, D, device=dev)
        with torch.no_grad():
            lp = flow.log_prob(x)
            assert lp.shape == (B,) and torch.isfinite(lp).all()
            assert flow.log_prob(x[:0]).shape == (0,)
            z, ld = flow.bijection.forward(x[:0])
            assert z.shape == (0, D) and ld.shape == (0,)
            xs, lps = flow.sample(B, return_log_prob=True)
            assert xs.shape == (B, D) and torch.isfinite(xs).all() and torch.isfinite(lps).all()
            # row i of a batch does not depend on the other rows or on the tile it lands in
            assert torch.equal(flow.log_prob(x[B // 2:B // 2 + 1]), lp[B // 2:B // 2 + 1])


@pytest.mark.parametrize('state', ['E', 'T'])
def test_mq128_many_tiles_vs_oracle(dev, state):
    """MaskedAutoregressiveRQNSF(128) through the persistent tcgen05 kernel at 32768 + 77 rows (257 tiles on 148 SMs: the
    mbarrier phases carry over tiles), states E and T, log_prob within 1e-4 abs/rel of the oracle (chunked on the CPU)."""
    import torchflows_b200.architectures as arch
    from torchflows_b200 import Flow, _native as N
    torch.manual_seed(3)
    D, B = 128, 32768 + 77
    flow = Flow(arch.MaskedAutoregressiveRQNSF(D)).to(dev)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, D, generator=g)
    if state == 'T':
        flow.train()
        with torch.no_grad():
            flow.log_prob((torch.randn(4096, D, generator=g) * 1.3 + 0.2).to(dev))
    flow.eval()
    o = fo.OracleFlow('MaskedAutoregressiveRQNSF', (D,), {k: v.cpu() for k, v in flow.state_dict().items()})
    with torch.no_grad():
        lp = flow.log_prob(x.to(dev))
    assert N.last_flow_kernel() == N.KERNEL_TCM
    ref = torch.cat([o.log_prob(x[i:i + 4096]) for i in range(0, B, 4096)])
    close(lp, ref, f'MQ128 log_prob state {state}', LP_TOL, LP_TOL)
