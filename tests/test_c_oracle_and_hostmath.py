"""CPU checks of (a) the plain-C oracle against the reference's golden vectors and (b) the kernels'
per-element math (host build of csrc/b2f_math.cuh) against the oracles."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, flow_oracle as fo
from tests import hostmath

TOL = 1e-5   # |d| <= 1e-5 + 1e-5 |ref|
LD_TOL = 1e-4  # north_star: per-sample log-density within 1e-4 abs/rel


def z_atol(boundary):
    """Two valid fp32 evaluations of the knots (2b*cumsum - b) differ by a few ulp of the boundary: the
    reference's own fp32 outputs are 2.5-3.0e-5 away from its fp64 outputs at boundary 50 (SURVEY Appendix C;
    ulp(50) = 3.8e-6, one per cumsum step that rounds differently).  Absolute tolerance on spline outputs:
    1e-5 + 16 ulp(boundary) (7.1e-5 at b=50, 1.8e-5 at b=5, 1.2e-5 at b=1)."""
    return 1e-5 + 16 * float(np.spacing(np.float32(boundary)))


def _close(a, b, what, atol=TOL, rtol=TOL):
    err = (a - b).abs() - (atol + rtol * b.abs())
    assert (err <= 0).all(), f'{what}: worst excess {err.max().item():.3e}, max abs diff {(a - b).abs().max().item():.3e}'


def _rq_cases(golden):
    return [c for c in golden('transformers.pt') if c['kind'] == 'rq']


def test_exp_det_accuracy_and_agreement():
    rng = np.random.default_rng(0)
    ts = np.concatenate([-rng.random(20000) * 86, -rng.random(20000) * 3, [0.0, -1e-8, -86.0, -100.0]]).astype(np.float32)
    worst = worst_top = 0.0
    for t in ts:
        a = c_oracle.exp_det(float(t))
        b = float(hostmath.lib().hm_exp_det(float(t)))
        assert a == b                                   # the two independent implementations agree bit for bit
        ref = np.exp(max(float(t), -86.0))
        ulp = abs(a - ref) / np.spacing(np.float32(ref))
        worst = max(worst, ulp)
        if t >= -3.0:
            worst_top = max(worst_top, ulp)
    assert worst_top < 1.05 and worst < 5.0      # the softmax terms that matter (within e^-3 of the maximum): < 1 ulp
    assert c_oracle.exp_det(0.0) == 1.0


def test_c_oracle_vs_reference_golden(golden):
    """Values within 1e-5; bin indices equal to the reference's except at (counted) ulp-level ties."""
    n_elem = n_flip = 0
    for c in _rq_cases(golden):
        for inverse, vin, vout, ldk, kk in ((False, 'x', 'z', 'ld_f', 'k_f'), (True, 'z', 'xr', 'ld_i', 'k_i')):
            out, ld, k = c_oracle.rq(c[vin], c['h'], c['n_bins'], c['boundary'], inverse)
            _close(out, c[vout], f"out nb={c['n_bins']} b={c['boundary']} inv={inverse}", atol=z_atol(c['boundary']))
            _close(ld.sum(-1), c[ldk], 'ld', atol=LD_TOL, rtol=LD_TOL)
            flips = (k.long() != c[kk])
            n_elem += k.numel()
            n_flip += int(flips.sum())
    assert n_flip <= max(1, n_elem // 100000), (n_flip, n_elem)
    for c in golden('transformers.pt'):
        if c['kind'] in ('affine', 'inverse_affine'):
            inv = c['kind'] == 'inverse_affine'
            out, ld = c_oracle.affine(c['x'], c['h'], inv)
            _close(out, c['z'], 'affine z')
            _close(ld.flatten(-len(c['event_shape'])).sum(-1), c['ld_f'], 'affine ld')


def test_hostmath_bins_bit_exact_vs_c_oracle(golden):
    """The kernel arithmetic (csrc/b2f_math.cuh) and the C oracle agree bit for bit on k."""
    for c in _rq_cases(golden):
        for inverse, vin in ((False, 'x'), (True, 'z')):
            for templated in (True, False):
                out, ld, k = hostmath.rq(c[vin], c['h'], c['n_bins'], c['boundary'], inverse, templated)
                o2, l2, k2 = c_oracle.rq(c[vin], c['h'], c['n_bins'], c['boundary'], inverse)
                assert torch.equal(k, k2)
                _close(out, o2, 'out', atol=z_atol(c['boundary']))
                _close(ld, l2, 'ld', atol=1e-5, rtol=1e-5)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(200000, generator=g) * 20
    h = torch.randn(200000, 23, generator=g) * 2
    for inverse in (False, True):
        out, ld, k = hostmath.rq(x, h, 8, 50.0, inverse)
        o2, l2, k2 = c_oracle.rq(x, h, 8, 50.0, inverse)
        assert torch.equal(k, k2)
        _close(out, o2, 'out big', atol=z_atol(50.0))


def test_hostmath_affine(golden):
    for c in golden('transformers.pt'):
        if c['kind'] in ('affine', 'inverse_affine'):
            inv = c['kind'] == 'inverse_affine'
            out, ld = hostmath.affine(c['x'], c['h'], inv)
            _close(out, c['z'], 'affine z')
            _close(ld.flatten(-len(c['event_shape'])).sum(-1), c['ld_f'], 'affine ld')


@pytest.mark.parametrize('n_bins', [8, 4])
def test_hostmath_rq_backward_vs_autograd(n_bins):
    g = torch.Generator().manual_seed(11)
    n = 4096
    x = (torch.randn(n, generator=g) * 3)
    x[:4] = torch.tensor([60.0, -70.0, 0.0, 49.9])
    h = torch.randn(n, 3 * n_bins - 1, generator=g)
    gz = torch.randn(n, generator=g)
    gl = torch.randn(n, generator=g)
    def autograd(dt):
        xd, hd = x.to(dt).requires_grad_(True), h.to(dt).requires_grad_(True)
        z, ld = fo.rq_forward(xd[:, None], hd[:, None, :], n_bins=n_bins, boundary=50.0)
        (z[:, 0] * gz.to(dt)).sum().add((ld * gl.to(dt)).sum()).backward()
        return xd.grad.double(), hd.grad.double()

    def rel(a, b):
        return ((a - b).norm() / b.norm()).item()

    x64, h64 = autograd(torch.float64)      # referee
    x32, h32 = autograd(torch.float32)      # what the fp32 reference's autograd produces
    dv, dh = hostmath.rq_backward(x, h, gz, gl, n_bins, 50.0)
    # The knot gradients cancel at the 1e-3 level in fp32 (reference fp32-vs-fp64: ~6e-4 relative L2 on dh),
    # so the bar is: no worse than the fp32 reference against the fp64 referee, and close to the fp32 reference.
    assert rel(dv.double(), x64) <= 2 * rel(x32, x64) + 1e-6
    assert rel(dh.double(), h64) <= 2 * rel(h32, h64) + 1e-6
    assert rel(dv.double(), x32) < 1e-5
    assert rel(dh.double(), h32) < 2e-3
    assert dv[0] == gz[0] and dv[1] == gz[1] and (dh[:2] == 0).all()      # out-of-bounds: dL/dv = GZ, dL/dh = 0


@pytest.mark.parametrize('boundary,scale', [(50.0, 3.0), (5.0, 2.0)])
def test_hostmath_fast_rq_backward_vs_autograd(boundary, scale):
    """rqf::backward_fwd (the backward epilogue of the wide-conditioner GEMM kernel, csrc/b2f_rqfast.cuh) against the
    reference's autograd through the oracle: same bar as rq_backward_fwd above."""
    g = torch.Generator().manual_seed(21)
    n = 8192
    x = torch.randn(n, generator=g) * scale
    x[:4] = torch.tensor([60.0, -70.0, 0.0, boundary - 0.1])
    h = torch.randn(n, 23, generator=g)
    gz = torch.randn(n, generator=g)
    gl = torch.randn(n, generator=g)

    def autograd(dt):
        xd, hd = x.to(dt).requires_grad_(True), h.to(dt).requires_grad_(True)
        z, ld = fo.rq_forward(xd[:, None], hd[:, None, :], n_bins=8, boundary=boundary)
        (z[:, 0] * gz.to(dt)).sum().add((ld * gl.to(dt)).sum()).backward()
        return xd.grad.double(), hd.grad.double()

    def rel(a, b):
        return ((a - b).norm() / b.norm()).item()

    x64, h64 = autograd(torch.float64)
    x32, h32 = autograd(torch.float32)
    dv, dh = hostmath.rq_backward_fast(x, h, gz, gl, boundary)
    assert torch.isfinite(dv).all() and torch.isfinite(dh).all()
    assert rel(dv.double(), x64) <= 2 * rel(x32, x64) + 1e-5
    assert rel(dh.double(), h64) <= 2 * rel(h32, h64) + 1e-5
    assert dv[0] == gz[0] and dv[1] == gz[1] and (dh[:2] == 0).all()
    # and against the deterministic-knot backward it replaces in that kernel
    dv0, dh0 = hostmath.rq_backward(x, h, gz, gl, 8, boundary)
    assert rel(dv.double(), dv0.double()) < 1e-5
    assert rel(dh.double(), dh0.double()) < 2e-3


@pytest.mark.parametrize('inverse', [False, True])
def test_hostmath_affine_backward_vs_autograd(inverse):
    g = torch.Generator().manual_seed(12)
    n = 2048
    x = torch.randn(n, generator=g)
    h = torch.randn(n, 2, generator=g)
    gz, gl = torch.randn(n, generator=g), torch.randn(n, generator=g)
    xd, hd = x.double().requires_grad_(True), h.double().requires_grad_(True)
    fn = fo.affine_inverse if inverse else fo.affine_forward
    z, ld = fn(xd[:, None], hd[:, None, :])
    (z[:, 0] * gz.double()).sum().add((ld * gl.double()).sum()).backward()
    dx, dh = hostmath.affine_backward(x, h, gz, gl, inverse)
    assert (dx.double() - xd.grad).norm() / xd.grad.norm() < 1e-5
    assert (dh.double() - hd.grad).norm() / hd.grad.norm() < 1e-5


@pytest.mark.parametrize('n_bins', [8, 4])
def test_hostmath_rq_inverse_backward_vs_autograd(n_bins):
    """Backward of the inverse-direction spline (implicit function theorem) against autograd through
    the oracle's rq_inverse."""
    g = torch.Generator().manual_seed(21)
    n = 4096
    z = torch.randn(n, generator=g) * 3
    z[:3] = torch.tensor([60.0, -70.0, 0.0])
    h = torch.randn(n, 3 * n_bins - 1, generator=g)
    gx, gl = torch.randn(n, generator=g), torch.randn(n, generator=g)

    def autograd(dt):
        zd, hd = z.to(dt).requires_grad_(True), h.to(dt).requires_grad_(True)
        x, ld = fo.rq_inverse(zd[:, None], hd[:, None, :], n_bins=n_bins, boundary=50.0)
        (x[:, 0] * gx.to(dt)).sum().add((ld * gl.to(dt)).sum()).backward()
        return zd.grad.double(), hd.grad.double()

    def rel(a, b):
        return ((a - b).norm() / b.norm()).item()

    z64, h64 = autograd(torch.float64)
    z32, h32 = autograd(torch.float32)
    dz, dh = hostmath.rq_backward_inv(z, h, gx, gl, n_bins, 50.0)
    assert rel(dz.double(), z64) <= 3 * rel(z32, z64) + 1e-5
    assert rel(dh.double(), h64) <= 3 * rel(h32, h64) + 1e-5
    assert dz[0] == gx[0] and (dh[:2] == 0).all()
