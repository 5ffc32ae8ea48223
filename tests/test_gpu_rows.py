"""Row-per-thread flow kernel (csrc/b2f_flow_rows.cu) against the generic kernel and the CPU oracle.

The default dispatch sends affine / shift programs with D % 8 == 0 and hidden width <= 31 to the rows kernel;
B2F_DISABLE_ROWS=1 + B2F_DISABLE_TC=1 forces the generic kernel (the environment is read at every call)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

@pytest.fixture(autouse=True)
def _without_the_multi_tile_affine_kernel():
    """These tests pin the dispatch of the older kernels; affine / shift coupling programs would otherwise go to the
    multi-tile tensor-core kernel (csrc/b2f_flow_tca.cu, covered by tests/test_gpu_tca.py)."""
    os.environ['B2F_DISABLE_TCA'] = '1'
    yield
    os.environ.pop('B2F_DISABLE_TCA', None)



def _run(flow, x, z):
    with torch.no_grad():
        lp = flow.log_prob(x)
        zf, ld = flow.bijection.forward(x)
        xi, ldi = flow.bijection.inverse(z)
        xs, lps = flow._sample_from_base(z, no_grad=True, return_log_prob=True)
    return lp, zf, ld, xi, ldi, xs, lps


def _generic(flow, x, z):
    os.environ['B2F_DISABLE_ROWS'] = '1'
    os.environ['B2F_DISABLE_TC'] = '1'
    try:
        return _run(flow, x, z)
    finally:
        os.environ.pop('B2F_DISABLE_ROWS', None)
        os.environ.pop('B2F_DISABLE_TC', None)


NAMES = ('log_prob', 'z', 'log_det', 'x_inv', 'log_det_inv', 'sample', 'sample log_prob')


@pytest.mark.parametrize('preset,D,B', [
    ('RealNVP', 64, 4096 + 77), ('RealNVP', 8, 1), ('RealNVP', 16, 31), ('RealNVP', 136, 33), ('InverseRealNVP', 32, 300),
    ('NICE', 64, 1111), ('NICE', 24, 129), ('MAF', 128, 700), ('MAF', 32, 64), ('MAF', 8, 5), ('IAF', 64, 200),
    ('IAF', 40, 1000)])
def test_rows_kernel_matches_generic_kernel_and_oracle(preset, D, B):
    from oracle.flow_oracle import OracleFlow
    from torchflows_b200 import Flow
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
    from torchflows_b200 import _native as N
    rows = _run(flow, x.to(dev), z.to(dev))
    assert N.last_flow_kernel() == N.KERNEL_ROWS
    gen = _generic(flow, x.to(dev), z.to(dev))
    assert N.last_flow_kernel() == N.KERNEL_GENERIC
    for a, b, n in zip(rows, gen, NAMES):
        a, b = a.double().cpu(), b.double().cpu()
        assert torch.isfinite(a).all(), n
        err = ((a - b).abs() / (1 + b.abs())).max().item()
        assert err < 2e-5, (n, err)
    nb = min(B, 256 if preset not in ('MAF', 'IAF') else 48)
    lp_ref = oracle.log_prob(x[:nb]).double()
    assert ((rows[0][:nb].double().cpu() - lp_ref).abs() / (1 + lp_ref.abs())).max().item() < 1e-4
    xs_ref, lps_ref = oracle.sample_from_noise(z[:nb], return_log_prob=True)
    assert ((rows[6][:nb].double().cpu() - lps_ref.double()).abs() / (1 + lps_ref.double().abs())).max().item() < 1e-4
    assert ((rows[5][:nb].double().cpu() - xs_ref.double()).abs() / (1 + xs_ref.double().abs())).max().item() < 1e-4


@pytest.mark.parametrize('tpw', [1, 3, 8])
def test_rows_kernel_tile_walk_and_ragged_tail(tpw):
    """Several tiles per warp, a batch that ends in the middle of a warp's tile, and a tail CTA with idle warps."""
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import RealNVP
    dev = torch.device('cuda:0')
    torch.manual_seed(3)
    flow = Flow(RealNVP(32)).to(dev).eval()
    for B in (1, 32, 33, 127, 128, 129, 128 * tpw + 5, 128 * tpw * 3 + 97):
        x, z = torch.randn(B, 32, device=dev), torch.randn(B, 32, device=dev)
        os.environ['B2F_ROWS_TPW'] = str(tpw)
        try:
            rows = _run(flow, x, z)
        finally:
            os.environ.pop('B2F_ROWS_TPW', None)
        gen = _generic(flow, x, z)
        for a, b, n in zip(rows, gen, NAMES):
            assert a.shape == b.shape
            err = ((a.double() - b.double()).abs() / (1 + b.double().abs())).max().item()
            assert err < 2e-5, (B, n, err)


def test_rows_kernel_raw_programs_through_the_c_abi():
    """Programs no preset produces: two elementwise runs separated only by a FLIP, a trailing FLIP, no elementwise layer at
    all, LOGP_OF_INPUT -- rows kernel vs generic kernel through b2f_flow_apply."""
    from torchflows_b200 import _native as N
    from torchflows_b200 import _program as P
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import RealNVP
    dev = torch.device('cuda:0')
    torch.manual_seed(11)
    D, B = 16, 777
    flow = Flow(RealNVP(D)).to(dev).eval()
    with torch.no_grad():
        for name, p in flow.named_parameters():
            if name.endswith('.value'):
                p.add_(0.3 * torch.randn_like(p))
    ops = flow.bijection.fused_ops('forward')
    assert ops is not None
    ew = [o for o in ops if o.kind == N.OP_ELEMENTWISE]
    fl = [o for o in ops if o.kind == N.OP_FLIP]
    cp = [o for o in ops if o.kind == N.OP_COUPLING]
    programs = [
        [ew[0], fl[0], ew[1], cp[0], fl[0]],
        [cp[0], fl[0], cp[1]],
        [fl[0], ew[0], ew[1], fl[0], fl[0], cp[1], ew[2], fl[0], ew[0]],
    ]
    x = torch.randn(B, D, device=dev)
    loc, lsc = 0.1 * torch.randn(D, device=dev), 0.1 * torch.randn(D, device=dev)
    for prog in programs:
        for flags in (0, N.FLOW_LOGP_OF_INPUT):
            outs = []
            for generic in (False, True):
                if generic:
                    os.environ['B2F_DISABLE_ROWS'] = '1'
                    os.environ['B2F_DISABLE_TC'] = '1'
                try:
                    with torch.no_grad():
                        outs.append(P.run_program(prog, x, want_log_prob=True, base_loc=loc, base_log_scale=lsc, flags=flags))
                finally:
                    os.environ.pop('B2F_DISABLE_ROWS', None)
                    os.environ.pop('B2F_DISABLE_TC', None)
            for a, b in zip(*outs):
                err = ((a.double() - b.double()).abs() / (1 + b.double().abs())).max().item()
                assert err < 2e-5, (len(prog), flags, err)


@pytest.mark.parametrize('preset,D,B', [('MaskedAutoregressiveRQNSF', 128, 333), ('MaskedAutoregressiveRQNSF', 16, 100),
                                        ('InverseAutoregressiveRQNSF', 24, 77), ('InverseAutoregressiveRQNSF', 64, 1000)])
def test_rows_kernel_sequential_spline_layers(preset, D, B):
    """The D-step sequential direction of spline MADE layers (MA-RQNSF sampling, IA-RQNSF density) on the rows kernel
    (per-warp streaming of the per-element output-layer weights), incl. the reference's last-iteration log-det, against
    the generic kernel and the CPU oracle."""
    from oracle.flow_oracle import OracleFlow
    from torchflows_b200 import Flow, _native as N
    import torchflows_b200.architectures as arch
    dev = torch.device('cuda:0')
    torch.manual_seed(D + 7)
    flow = Flow(getattr(arch, preset)(D)).eval()
    with torch.no_grad():
        for name, p in flow.named_parameters():
            if name.endswith('.value'):
                p.add_(0.3 * torch.randn_like(p))
    oracle = OracleFlow(preset, (D,), flow.state_dict())
    flow = flow.to(dev)
    g = torch.Generator().manual_seed(B)
    v = (1.5 * torch.randn(B, D, generator=g)).to(dev)
    sampling = preset == 'MaskedAutoregressiveRQNSF'

    def call():
        with torch.no_grad():
            if sampling:
                return flow._sample_from_base(v, no_grad=True, return_log_prob=True)
            z, ld = flow.bijection.forward(v)
            return z, flow.log_prob(v), ld
    rows = call()
    assert N.last_flow_kernel() == N.KERNEL_ROWS
    os.environ['B2F_DISABLE_ROWS'] = '1'
    try:
        gen = call()
        assert N.last_flow_kernel() == N.KERNEL_GENERIC
    finally:
        os.environ.pop('B2F_DISABLE_ROWS', None)
    # two fp32 kernels with different summation orders in the hidden layer: spline VALUES agree to the spline tolerance
    # (knot noise ~ ulp(boundary), cf. tests/test_gpu_tc.py), log-quantities to 2e-4
    for idx, (a, b) in enumerate(zip(rows, gen)):
        err = ((a.double() - b.double()).abs() / (1 + b.double().abs())).max().item()
        assert err < (2e-3 if idx == 0 else 2e-4), (idx, err)
    nb = min(B, 48)
    if sampling:
        xs_ref, lps_ref = oracle.sample_from_noise(v[:nb].cpu(), return_log_prob=True)
        assert ((rows[0][:nb].double().cpu() - xs_ref.double()).abs() / (1 + xs_ref.double().abs())).max().item() < 2e-3
        assert ((rows[1][:nb].double().cpu() - lps_ref.double()).abs() / (1 + lps_ref.double().abs())).max().item() < 2e-4
    else:
        lp_ref = oracle.log_prob(v[:nb].cpu()).double()
        assert ((rows[1][:nb].double().cpu() - lp_ref).abs() / (1 + lp_ref.abs())).max().item() < 1e-4


def test_rows_kernel_in_place_spline_sampling_is_bit_identical():
    """MA-RQNSF sampling keeps its rows in the output buffer (global memory, no shared-memory tile); same arithmetic as
    the tiled mode (B2F_ROWS_NO_INPLACE=1), so bit-identical results, including a ragged last warp."""
    from torchflows_b200 import Flow, _native as N
    from torchflows_b200.architectures import MaskedAutoregressiveRQNSF
    dev = torch.device('cuda:0')
    torch.manual_seed(4)
    flow = Flow(MaskedAutoregressiveRQNSF(32)).to(dev).eval()
    with torch.no_grad():
        for name, p in flow.named_parameters():
            if name.endswith('.value'):
                p.add_(0.3 * torch.randn_like(p))
    for B in (1, 37, 64, 1000):
        z = 1.5 * torch.randn(B, 32, device=dev)
        outs = []
        for tiled in (False, True):
            if tiled:
                os.environ['B2F_ROWS_NO_INPLACE'] = '1'
            try:
                with torch.no_grad():
                    outs.append(flow._sample_from_base(z, no_grad=True, return_log_prob=True))
                assert N.last_flow_kernel() == N.KERNEL_ROWS
            finally:
                os.environ.pop('B2F_ROWS_NO_INPLACE', None)
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), B


def test_sequential_spline_folded_operands_agree_with_the_plain_ones():
    """MA-RQNSF sampling on the rows kernel with the folded output layer (B2F_FLAG_SEQ_FOLDED, csrc/b2f_rqfast.cuh
    sequential_step) against the same kernel fed the plain 23-parameter operands (B2F_NO_SEQ_FOLD=1), incl. the reference's
    last-iteration log-det and its exact variant."""
    from torchflows_b200 import Flow, _native as N
    import torchflows_b200.architectures as arch
    dev = torch.device('cuda:0')
    torch.manual_seed(5)
    flow = Flow(arch.MaskedAutoregressiveRQNSF(64)).to(dev).eval()
    with torch.no_grad():
        for name, p in flow.named_parameters():
            if name.endswith('.value'):
                p.add_(0.3 * torch.randn_like(p))
    z = 1.5 * torch.randn(3000, 64, device=dev)
    for quirk in (True, False):
        for layer in flow.bijection.layers:
            if hasattr(layer, 'sequential_log_det_reference_quirk'):
                layer.sequential_log_det_reference_quirk = quirk
        with torch.no_grad():
            xs, lps = flow._sample_from_base(z, no_grad=True, return_log_prob=True)
            assert N.last_flow_kernel() == N.KERNEL_ROWS
            os.environ['B2F_NO_SEQ_FOLD'] = '1'
            try:
                xs0, lps0 = flow._sample_from_base(z, no_grad=True, return_log_prob=True)
                assert N.last_flow_kernel() == N.KERNEL_ROWS
            finally:
                os.environ.pop('B2F_NO_SEQ_FOLD', None)
        assert not torch.equal(xs, xs0)          # two different formulations really ran
        err = ((xs.double() - xs0.double()).abs() / (1 + xs0.double().abs())).max().item()
        err_lp = ((lps.double() - lps0.double()).abs() / (1 + lps0.double().abs())).max().item()
        assert err < 2e-3 and err_lp < 1e-4, (quirk, err, err_lp)


def test_folded_sequential_operands_fall_back_when_the_rows_kernel_declines():
    """A 4-byte-aligned input makes the rows kernel decline the program; the library then rejects the folded operands
    (only the rows kernel takes them) and the host retries with the plain ones on the generic kernel: same samples."""
    from torchflows_b200 import Flow, _native as N
    import torchflows_b200.architectures as arch
    dev = torch.device('cuda:0')
    torch.manual_seed(9)
    D, B = 16, 200
    flow = Flow(arch.MaskedAutoregressiveRQNSF(D)).to(dev).eval()
    storage = torch.randn(B * D + 1, device=dev)
    z_misaligned = storage[1:].view(B, D)
    assert z_misaligned.data_ptr() % 16 != 0
    z = z_misaligned.clone()
    with torch.no_grad():
        xs = flow._sample_from_base(z, no_grad=True)
        assert N.last_flow_kernel() == N.KERNEL_ROWS
        xs_m = flow._sample_from_base(z_misaligned, no_grad=True)
        assert N.last_flow_kernel() == N.KERNEL_GENERIC
    err = ((xs.double() - xs_m.double()).abs() / (1 + xs.double().abs())).max().item()
    assert err < 2e-3, err
    # the layers remember: no second failed launch
    assert all(getattr(l, '_b2f_no_seq_fold', False) for l in flow.bijection.layers if hasattr(l, 'sequential_log_det_reference_quirk'))
