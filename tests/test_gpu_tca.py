"""Multi-tile affine / shift coupling kernel (csrc/b2f_flow_tca.cu) against the generic fp32 kernel and the CPU oracle
(oracle/flow_oracle.py, the reference's ATen ops: architectures.py:57-96, layers_base.py:119-163, affine.py:33-59,149-159).
The conditioner runs as 3xTF32 on the tensor cores and must be fp32-faithful: same tolerances as the
row-per-thread kernel's tests (values 1e-5, log-quantities 1e-4 abs/rel -- the north-star bar)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu



@pytest.fixture(autouse=True)
def _every_shape_that_fits():
    """By default only event sizes with four tile pipelines per SM (D <= 64) go to this kernel; the tests also cover the
    wider shapes it can run."""
    os.environ['B2F_TCA_ANY'] = '1'
    yield
    os.environ.pop('B2F_TCA_ANY', None)


NAMES = ('log_prob', 'z', 'log_det', 'x_inv', 'log_det_inv', 'sample', 'sample log_prob')


def _run(flow, x, z):
    with torch.no_grad():
        lp = flow.log_prob(x)
        zz, ld = flow.bijection.forward(x)
        xi, ldi = flow.bijection.inverse(z)
        xs, lps = flow._sample_from_base(z, no_grad=True, return_log_prob=True)
    return lp, zz, ld, xi, ldi, xs, lps


def _generic(flow, x, z):
    os.environ['B2F_DISABLE_ROWS'] = '1'
    os.environ['B2F_DISABLE_TC'] = '1'
    try:
        return _run(flow, x, z)
    finally:
        os.environ.pop('B2F_DISABLE_ROWS', None)
        os.environ.pop('B2F_DISABLE_TC', None)


@pytest.mark.parametrize('preset,D,B', [
    ('RealNVP', 64, 4096 + 77), ('RealNVP', 64, 1), ('RealNVP', 32, 300), ('RealNVP', 128, 1000), ('RealNVP', 96, 515),
    ('InverseRealNVP', 32, 300), ('NICE', 64, 1111), ('NICE', 32, 129), ('RealNVP', 64, 148 * 4 * 128 * 2 + 333)])
def test_tca_kernel_matches_generic_kernel_and_oracle(preset, D, B):
    from oracle.flow_oracle import OracleFlow
    from torchflows_b200 import Flow, _native as N
    import torchflows_b200.architectures as arch
    dev = torch.device('cuda:0')
    torch.manual_seed(D + 1)
    flow = Flow(getattr(arch, preset)(D)).eval()
    with torch.no_grad():                       # random-init elementwise layers are the identity: make them count
        for name, p in flow.named_parameters():
            if name.endswith('.value'):
                p.add_(0.3 * torch.randn_like(p))
    oracle = OracleFlow(preset, (D,), flow.state_dict())
    flow = flow.to(dev)
    g = torch.Generator().manual_seed(B)
    x, z = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    ours = _run(flow, x.to(dev), z.to(dev))
    assert N.last_flow_kernel() == N.KERNEL_TCA
    gen = _generic(flow, x.to(dev), z.to(dev))
    assert N.last_flow_kernel() == N.KERNEL_GENERIC
    # fp64 referee: the 3xTF32 products carry ~2^-21 per term (plain FFMA: 2^-24), so on an ill-conditioned column both
    # kernels lose digits; ours may lose at most 4x what the fp32 FFMA kernel loses, floor 2e-5 (values) / 1e-4 (log-quantities)
    o64 = OracleFlow(preset, (D,), {k: v.double() for k, v in flow.state_dict().items()})
    with torch.no_grad():
        z64, ld64 = o64.forward(x.double())
        xi64, ldi64 = o64.inverse(z.double())
        xs64, lps64 = o64.sample_from_noise(z.double(), return_log_prob=True)
        ref64 = (o64.log_prob(x.double()), z64, ld64, xi64, ldi64, xs64, lps64)
    for a, b, r, n in zip(ours, gen, ref64, NAMES):
        a, b = a.double().cpu(), b.double().cpu()
        assert torch.isfinite(a).all(), n
        err = ((a - r).abs() / (1 + r.abs())).max().item()
        err_generic = ((b - r).abs() / (1 + r.abs())).max().item()
        assert err < max(1e-4 if 'log' in n else 2e-5, 4 * err_generic), (n, err, err_generic)
    nb = min(B, 512)
    with torch.no_grad():
        lp_ref = oracle.log_prob(x[:nb]).double()
        xs_ref, lps_ref = oracle.sample_from_noise(z[:nb], return_log_prob=True)
    assert ((ours[0][:nb].double().cpu() - lp_ref).abs() / (1 + lp_ref.abs())).max().item() < 1e-4
    assert ((ours[5][:nb].double().cpu() - xs_ref.double()).abs() / (1 + xs_ref.double().abs())).max().item() < 2e-5
    assert ((ours[6][:nb].double().cpu() - lps_ref.double()).abs() / (1 + lps_ref.double().abs())).max().item() < 1e-4


def test_tca_kernel_deeper_flows_and_a_non_standard_base():
    from oracle.flow_oracle import OracleFlow
    from torchflows_b200 import Flow, _native as N
    from torchflows_b200.architectures import RealNVP
    dev = torch.device('cuda:0')
    torch.manual_seed(9)
    flow = Flow(RealNVP(64, n_layers=4)).eval()
    with torch.no_grad():
        for name, p in flow.named_parameters():
            if name.endswith('.value'):
                p.add_(0.2 * torch.randn_like(p))
    oracle = OracleFlow('RealNVP', (64,), flow.state_dict(), n_layers=4)
    flow = flow.to(dev)
    x = torch.randn(2000, 64)
    with torch.no_grad():
        lp = flow.log_prob(x.to(dev)).cpu().double()
        assert N.last_flow_kernel() == N.KERNEL_TCA
        ref = oracle.log_prob(x).double()
    assert ((lp - ref).abs() / (1 + ref.abs())).max().item() < 1e-4


def test_tca_sample_draws_noise_in_the_kernel():
    """Flow.sample of an affine coupling flow: same stream, same samples as the materialised-noise path."""
    from torchflows_b200 import Flow, _native as N, _program as prog
    from torchflows_b200.architectures import RealNVP
    dev = torch.device('cuda:0')
    torch.manual_seed(2)
    flow = Flow(RealNVP(64)).to(dev).eval()
    with torch.no_grad():
        torch.manual_seed(21)
        x1, lp1 = flow.sample((5000,), no_grad=True, return_log_prob=True)
        assert N.last_flow_kernel() == N.KERNEL_TCA
        torch.manual_seed(21)
        seed, offset = prog.next_noise_stream()
        z = N.philox_normal(5000, 64, dev, seed, offset)
        x2, lp2 = flow._sample_from_base(z, no_grad=True, return_log_prob=True)
    assert float((x1 - x2).abs().max()) <= 1e-5 * (1 + float(x2.abs().max()))
    assert float((lp1 - lp2).abs().max()) <= 1e-4 * (1 + float(lp2.abs().max()))


def test_tca_sample_with_a_non_standard_base():
    from torchflows_b200 import Flow, _native as N, _program as prog
    from torchflows_b200.architectures import RealNVP
    dev = torch.device('cuda:0')
    torch.manual_seed(6)
    flow = Flow(RealNVP(32)).to(dev).eval()
    with torch.no_grad():
        flow.base.loc.copy_(torch.randn(32, device=dev))
        flow.base.log_scale.copy_(torch.randn(32, device=dev) * 0.2)
        torch.manual_seed(9)
        x, lp = flow.sample((3000,), no_grad=True, return_log_prob=True)
        assert N.last_flow_kernel() == N.KERNEL_TCA
        torch.manual_seed(9)
        seed, offset = prog.next_noise_stream()
        z = N.philox_normal(3000, 32, dev, seed, offset, flow.base.loc.reshape(-1), flow.base.log_scale.reshape(-1))
        x2, lp2 = flow._sample_from_base(z, no_grad=True, return_log_prob=True)
    assert float((x - x2).abs().max()) <= 1e-5 * (1 + float(x2.abs().max()))
    assert float((lp - lp2).abs().max()) <= 1e-4 * (1 + float(lp2.abs().max()))
