"""One-pass MADE spline kernel (csrc/b2f_flow_tcm.cu) against the CPU oracle: MaskedAutoregressiveRQNSF densities and
InverseAutoregressiveRQNSF sampling passes, more tiles than SMs, ragged last tile, TMA and manual tile IO, weight states
E and T, out-of-bounds inputs, in-kernel noise.  Tolerance: the north star's 1e-4 abs/rel on log_prob."""
import os

import pytest
import torch

from oracle.flow_oracle import OracleFlow

pytestmark = pytest.mark.gpu
LP_TOL = 1e-4


def close(a, b, what, atol, rtol):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.isfinite(a).all(), what
    err = (a - b).abs() - (atol + rtol * b.abs())
    assert (err <= 0).all(), f'{what}: max abs diff {(a - b).abs().max().item():.3e}, worst excess {err.max().item():.3e}'


def make(preset, D, state, dev, n_layers=2):
    from torchflows_b200 import Flow
    import torchflows_b200.architectures as arch
    torch.manual_seed(200 + D)
    flow = Flow(getattr(arch, preset)(D, n_layers=n_layers)).to(dev)
    if state == 'T':                    # ActNorm data-initialised by one training-mode pass (state T)
        g = torch.Generator().manual_seed(5)
        flow.train()
        with torch.no_grad():
            flow.log_prob((torch.randn(2048, D, generator=g) * 1.3 + 0.2).to(dev))
    flow.eval()
    sd = {k: v.detach().cpu() for k, v in flow.state_dict().items()}
    return flow, OracleFlow(preset, (D,), sd, n_layers=n_layers)


def chunked(fn, x, chunk=4096):
    outs = [fn(x[i:i + chunk]) for i in range(0, x.shape[0], chunk)]
    if isinstance(outs[0], tuple):
        return tuple(torch.cat([o[j] for o in outs]) for j in range(len(outs[0])))
    return torch.cat(outs)


@pytest.mark.parametrize('tma', [True, False])
@pytest.mark.parametrize('D,B,state', [(128, 1000, 'E'), (128, 128 * 150 + 77, 'T'), (64, 4096 + 5, 'E'), (32, 5, 'E'),
                                       (96, 777, 'T')])
def test_density_direction_matches_oracle(D, B, state, tma):
    from torchflows_b200 import _native as N
    dev = torch.device('cuda:0')
    flow, oracle = make('MaskedAutoregressiveRQNSF', D, state, dev)
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, D, generator=g) * 1.5
    if tma:
        os.environ.pop('B2F_TCM_NO_TMA', None)
    else:
        os.environ['B2F_TCM_NO_TMA'] = '1'
    try:
        with torch.no_grad():
            lp = flow.log_prob(x.to(dev))
            assert N.last_flow_kernel() == N.KERNEL_TCM
            zf, ld = flow.bijection.forward(x.to(dev))
            assert N.last_flow_kernel() == N.KERNEL_TCM
            torch.cuda.synchronize()
    finally:
        os.environ.pop('B2F_TCM_NO_TMA', None)
    z_ref, ld_ref = chunked(oracle.forward, x)
    close(lp, chunked(oracle.log_prob, x), 'log_prob', LP_TOL, LP_TOL)
    close(ld, ld_ref, 'log_det', LP_TOL, LP_TOL)
    close(zf, z_ref, 'z', 1e-3, 1e-4)           # spline values behind a TF32 conditioner over all D inputs


@pytest.mark.parametrize('D,B,state', [(128, 128 * 160 + 3, 'E'), (64, 999, 'T')])
def test_sampling_direction_matches_oracle(D, B, state):
    """InverseAutoregressiveRQNSF: the sampling direction is the one-pass one (layers_base.py:202-211 through the inverse
    wrapper), with the base density of the input rows for return_log_prob=True (flows.py:710-712)."""
    from torchflows_b200 import _native as N
    dev = torch.device('cuda:0')
    flow, oracle = make('InverseAutoregressiveRQNSF', D, state, dev)
    g = torch.Generator().manual_seed(B)
    z = torch.randn(B, D, generator=g)
    with torch.no_grad():
        xs, lps = flow._sample_from_base(z.to(dev), no_grad=True, return_log_prob=True)
        assert N.last_flow_kernel() == N.KERNEL_TCM
        xs2 = flow._sample_from_base(z.to(dev), no_grad=True)
    xs_ref, lps_ref = chunked(lambda t: oracle.sample_from_noise(t, return_log_prob=True), z)
    close(xs, xs_ref, 'sample', 2e-3, 1e-4)
    close(lps, lps_ref, 'sample log_prob', 2 * LP_TOL, 2 * LP_TOL)
    assert torch.equal(xs, xs2)


def test_four_layers_and_out_of_bounds():
    from torchflows_b200 import _native as N
    dev = torch.device('cuda:0')
    flow, oracle = make('MaskedAutoregressiveRQNSF', 64, 'E', dev, n_layers=4)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3000, 64, generator=g) * 30.0
    with torch.no_grad():
        lp = flow.log_prob(x.to(dev))
        zf, ld = flow.bijection.forward(x.to(dev))
    assert N.last_flow_kernel() == N.KERNEL_TCM
    close(lp, oracle.log_prob(x), 'log_prob', LP_TOL, LP_TOL)
    z_ref, ld_ref = oracle.forward(x)
    close(zf, z_ref, 'z', 2e-3, 1e-3)
    close(ld, ld_ref, 'log_det', LP_TOL, LP_TOL)


def test_agrees_with_the_first_generation_kernel():
    """Same program through csrc/b2f_flow_tc.cu (B2F_DISABLE_TCM): two independent implementations of the path."""
    from torchflows_b200 import _native as N
    dev = torch.device('cuda:0')
    flow, _ = make('MaskedAutoregressiveRQNSF', 128, 'T', dev)
    x = torch.randn(20000, 128, device=dev)
    with torch.no_grad():
        lp = flow.log_prob(x)
        assert N.last_flow_kernel() == N.KERNEL_TCM
        os.environ['B2F_DISABLE_TCM'] = '1'
        try:
            lp0 = flow.log_prob(x)
            assert N.last_flow_kernel() != N.KERNEL_TCM
        finally:
            os.environ.pop('B2F_DISABLE_TCM', None)
    close(lp, lp0, 'log_prob', 2 * LP_TOL, 2 * LP_TOL)


def test_library_noise_sampling_is_deterministic_and_the_same_stream_as_the_materialised_one():
    """Flow.sample of InverseAutoregressiveRQNSF draws its noise inside the launch (Philox): same seed -> same rows, and the
    same rows (to kernel tolerance) as the first-generation kernel fed the materialised stream (b2f_philox_normal)."""
    from torchflows_b200 import _native as N
    dev = torch.device('cuda:0')
    flow, _ = make('InverseAutoregressiveRQNSF', 64, 'E', dev)
    torch.manual_seed(7)
    a, lpa = flow.sample(5000, no_grad=True, return_log_prob=True)
    assert N.last_flow_kernel() == N.KERNEL_TCM
    torch.manual_seed(7)
    b = flow.sample(5000, no_grad=True)
    assert torch.equal(a, b)
    os.environ['B2F_DISABLE_TCM'] = '1'
    try:
        torch.manual_seed(7)
        c, lpc = flow.sample(5000, no_grad=True, return_log_prob=True)
        assert N.last_flow_kernel() != N.KERNEL_TCM
    finally:
        os.environ.pop('B2F_DISABLE_TCM', None)
    close(a, c, 'sample', 2e-3, 1e-4)
    close(lpa, lpc, 'sample log_prob', 2 * LP_TOL, 2 * LP_TOL)
