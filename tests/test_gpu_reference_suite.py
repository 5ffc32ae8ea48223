"""The reference's own hot-path test files re-pointed at torchflows_b200 (SURVEY section 4 / check P6), run on the GPU.
Same parametrisations, seeds and tolerances as /root/reference/test (constants.py: data / log-det atol 1e-3, "easy" 1e-2);
context-conditioned cases are the ones not ported (context is out of this round's scope)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

BATCH_SHAPES = [(1,), (2,), (5,), (5, 2, 3)]           # test/constants.py:2
EVENT_SHAPES = [(2,), (3,), (3, 5, 2)]                 # test/constants.py:3
ATOL = 1e-3                                            # test/constants.py:11-13
DEV = 'cuda:0'


def _reconstruction(bijection, x, eps=ATOL):
    """test/test_reconstruction_bijections.py:64-93."""
    torch.manual_seed(0)
    z, ld_f = bijection.forward(x)
    xr, ld_i = bijection.inverse(z)
    batch_shape = x.shape[:x.dim() - len(bijection.event_shape)]
    assert x.shape == z.shape and ld_f.shape == ld_i.shape == batch_shape
    for t in (z, xr, ld_f, ld_i):
        assert t.isfinite().all()
    assert torch.allclose(x, xr, atol=eps), f'E: {(x - xr).abs().max():.3e}'
    assert torch.allclose(ld_f, -ld_i, atol=eps), f'E: {(ld_f + ld_i).abs().max():.3e}'


def _presets():
    from torchflows_b200.bijections.finite.autoregressive import architectures as a
    return a


@pytest.mark.parametrize('name', ['ReversePermutationMatrix', 'ElementwiseAffine', 'ElementwiseShift', 'ActNorm'])
@pytest.mark.parametrize('batch_shape', BATCH_SHAPES)
@pytest.mark.parametrize('event_shape', EVENT_SHAPES)
def test_linear(name, batch_shape, event_shape):
    """test/test_reconstruction_bijections.py:126-143 (hot-path subset)."""
    from torchflows_b200.bijections.finite.autoregressive import layers
    from torchflows_b200.bijections.finite.matrix import ReversePermutationMatrix
    cls = ReversePermutationMatrix if name == 'ReversePermutationMatrix' else getattr(layers, name)
    torch.manual_seed(0)
    x = torch.randn(*batch_shape, *event_shape, device=DEV)
    _reconstruction(cls(event_shape, context_shape=None).to(DEV), x)


@pytest.mark.parametrize('name', ['NICE', 'RealNVP', 'CouplingRQNSF'])
@pytest.mark.parametrize('batch_shape', BATCH_SHAPES)
@pytest.mark.parametrize('event_shape', EVENT_SHAPES)
def test_coupling(name, batch_shape, event_shape):
    """test/test_reconstruction_bijections.py:146-162."""
    torch.manual_seed(0)
    x = torch.randn(*batch_shape, *event_shape, device=DEV)
    _reconstruction(getattr(_presets(), name)(event_shape, context_shape=None).to(DEV), x)


@pytest.mark.parametrize('name', ['MAF', 'IAF', 'InverseAutoregressiveRQNSF', 'MaskedAutoregressiveRQNSF'])
@pytest.mark.parametrize('batch_shape', BATCH_SHAPES)
@pytest.mark.parametrize('event_shape', EVENT_SHAPES)
def test_masked_autoregressive(name, batch_shape, event_shape):
    """test/test_reconstruction_bijections.py:165-178."""
    torch.manual_seed(0)
    x = torch.randn(*batch_shape, *event_shape, device=DEV)
    _reconstruction(getattr(_presets(), name)(event_shape, context_shape=None).to(DEV), x)


@pytest.mark.parametrize('name', ['AffineCoupling', 'RQSCoupling', 'InverseAffineCoupling', 'ShiftCoupling',
                                  'AffineForwardMaskedAutoregressive', 'AffineInverseMaskedAutoregressive',
                                  'ElementwiseAffine', 'ElementwiseRQSpline', 'ElementwiseShift',
                                  'RQSForwardMaskedAutoregressive', 'RQSInverseMaskedAutoregressive'])
def test_identity_at_zero_parameters(name):
    """test/test_identity_bijections.py:56-68."""
    from torchflows_b200.bijections.finite.autoregressive import layers
    torch.manual_seed(0)
    x = torch.randn(2, 3, device=DEV)
    layer = getattr(layers, name)(event_shape=torch.Size((3,))).to(DEV)
    with torch.no_grad():
        for p in layer.parameters():
            p.data *= 0
    assert torch.allclose(layer(x)[0], x, atol=1e-2)
    assert torch.allclose(layer.inverse(x)[0], x, atol=1e-2)


@pytest.mark.parametrize('name', ['Affine', 'Shift', 'RationalQuadratic'])
@pytest.mark.parametrize('batch_shape', BATCH_SHAPES)
@pytest.mark.parametrize('event_shape', EVENT_SHAPES)
def test_transformers(name, batch_shape, event_shape):
    """test/test_reconstruction_transformers.py:22-58,68-84."""
    from torchflows_b200.bijections.finite.autoregressive.transformers.linear.affine import Affine, Shift
    from torchflows_b200.bijections.finite.autoregressive.transformers.spline.rational_quadratic import RationalQuadratic
    tr = {'Affine': Affine, 'Shift': Shift, 'RationalQuadratic': RationalQuadratic}[name](event_shape)
    torch.manual_seed(0)
    x = torch.randn(*batch_shape, *event_shape, device=DEV)
    h = torch.randn(*batch_shape, *tr.parameter_shape, device=DEV)
    z, ld_f = tr.forward(x, h)
    xr, ld_i = tr.inverse(z, h)
    assert z.shape == x.shape and ld_f.shape == ld_i.shape == batch_shape
    assert torch.allclose(x, xr, atol=ATOL) and torch.allclose(ld_f, -ld_i, atol=ATOL)


def test_spline_1d_2d_and_outside_boundary():
    """test/test_spline.py:30-79: the 2-D case has one input outside boundary=5 (identity tail)."""
    from torchflows_b200.bijections.finite.autoregressive.transformers.spline.rational_quadratic import RationalQuadratic
    torch.manual_seed(0)
    for event_shape, xs in (((1,), [[1.2], [4.0], [-3.6]]), ((2,), [[1.2, 2.0], [4.0, 6.0], [-3.6, 0.7]])):
        spline = RationalQuadratic(event_shape=event_shape, n_bins=8, boundary=5.0)
        x = torch.tensor(xs, device=DEV)
        h = torch.randn(3, *spline.parameter_shape, device=DEV)
        z, ld = spline(x, h)
        xr, ldi = spline.inverse(z, h)
        assert torch.allclose(x, xr, atol=ATOL) and torch.allclose(ld, -ldi, atol=ATOL)
    assert z[1, 1] == 6.0


@pytest.mark.parametrize('boundary', [1.0, 5.0, 50.0])
@pytest.mark.parametrize('batch_shape', BATCH_SHAPES)
@pytest.mark.parametrize('event_shape', EVENT_SHAPES)
def test_spline_exhaustive(boundary, batch_shape, event_shape):
    """test/test_spline.py:92-109."""
    from torchflows_b200.bijections.finite.autoregressive.transformers.spline.rational_quadratic import RationalQuadratic
    torch.manual_seed(0)
    spline = RationalQuadratic(event_shape=event_shape, n_bins=8, boundary=boundary)
    x = torch.randn(*batch_shape, *event_shape, device=DEV)
    h = torch.randn(*batch_shape, *spline.parameter_shape, device=DEV)
    z, ld = spline(x, h)
    xr, ldi = spline.inverse(z, h)
    assert z.shape == x.shape and ld.shape == batch_shape
    assert torch.allclose(x, xr, atol=ATOL) and torch.allclose(ld, -ldi, atol=ATOL)


@pytest.mark.parametrize('n_bins', [2, 4, 8, 16, 32])
@pytest.mark.parametrize('scale', [1e-2, 1.0, 1e2])
def test_rq_spline_bins_and_scales(n_bins, scale):
    """test/test_spline.py:116-135 (n_data, n_dim in {1,2,5,100,500} collapsed to two shapes): log-det antisymmetry."""
    from torchflows_b200.bijections.finite.autoregressive.transformers.spline.rational_quadratic import RationalQuadratic
    torch.manual_seed(0)
    for n_data, n_dim in ((5, 2), (100, 500)):
        spline = RationalQuadratic(event_shape=(n_dim,), n_bins=n_bins)
        x = torch.randn(n_data, n_dim, device=DEV) * scale
        h = torch.randn(n_data, n_dim, 3 * n_bins - 1, device=DEV)
        z, ld = spline.forward(x, h)
        xr, ldi = spline.inverse(z, h)
        assert torch.allclose(ld, -ldi, atol=1e-2 if n_dim == 500 else ATOL)


@pytest.mark.parametrize('name', ['ElementwiseAffine', 'AffineCoupling'])
def test_layer_gradients_image_event(name):
    """test/test_layer_gradients.py:12-31: loss.backward() through layers with event shape (3, 20, 20)."""
    from torchflows_b200.bijections.finite.autoregressive import layers
    torch.manual_seed(0)
    event_shape = (3, 20, 20)
    layer = getattr(layers, name)(event_shape).to(DEV)
    x = torch.randn(4, *event_shape, device=DEV)
    z, ld = layer(x)
    (z.sum() + ld.sum()).backward()
    grads = [p.grad for p in layer.parameters() if p.requires_grad and p.numel()]
    assert grads and all(g is not None and g.isfinite().all() for g in grads)


@pytest.mark.parametrize('name', ['RealNVP', 'MAF', 'CouplingRQNSF', 'MaskedAutoregressiveRQNSF', 'NICE', 'IAF'])
def test_deepcopy_and_cuda_inputs(name):
    """test/test_deepcopy.py:13-36 and test/test_cuda.py:7-64: deepcopy works; CPU and GPU data are accepted by
    log_prob and fit of a .cuda() flow."""
    from copy import deepcopy
    from torchflows_b200 import Flow
    torch.manual_seed(0)
    batch_shape, event_shape = (3, 5), (7, 11)
    x = torch.randn(*batch_shape, *event_shape)
    flow = Flow(getattr(_presets(), name)(event_shape)).cuda()
    deepcopy(flow)
    a = flow.log_prob(x)
    b = flow.log_prob(x.cuda())
    assert a.shape == batch_shape and torch.allclose(a, b)
    if True:
        flow.fit(x.reshape(-1, *event_shape), n_epochs=3)
        flow.fit(x.reshape(-1, *event_shape).cuda(), n_epochs=3)
        deepcopy(flow)


# ---------------------------------------------------------------------------------------------------------------------
# context-conditioned parametrisations (test/constants.py:5), SURVEY section 8f-1
# ---------------------------------------------------------------------------------------------------------------------
CONTEXT_SHAPES = [(2,), (3,), (3, 5, 2)]


@pytest.mark.parametrize('name', ['NICE', 'RealNVP', 'CouplingRQNSF', 'MAF', 'IAF', 'InverseAutoregressiveRQNSF',
                                  'MaskedAutoregressiveRQNSF'])
@pytest.mark.parametrize('batch_shape', [(2,), (5, 2, 3)])
@pytest.mark.parametrize('event_shape', EVENT_SHAPES)
@pytest.mark.parametrize('context_shape', CONTEXT_SHAPES)
def test_presets_with_context(name, batch_shape, event_shape, context_shape):
    """test/test_reconstruction_bijections.py:146-178 with context and test/test_autograd_bijections.py:41-54."""
    from torchflows_b200 import Flow
    torch.manual_seed(0)
    x = torch.randn(*batch_shape, *event_shape, device=DEV)
    c = torch.randn(*batch_shape, *context_shape, device=DEV)
    bij = getattr(_presets(), name)(event_shape, context_shape=context_shape).to(DEV)
    z, ld_f = bij.forward(x, context=c)
    xr, ld_i = bij.inverse(z, context=c)
    assert z.shape == x.shape and ld_f.shape == ld_i.shape == batch_shape
    assert torch.allclose(x, xr, atol=ATOL) and torch.allclose(ld_f, -ld_i, atol=ATOL)
    if name != 'InverseAutoregressiveRQNSF':      # spline + sequential density: gradient only with the exact log-det flag
        xc = x.clone().requires_grad_(True)
        lp = Flow(bij).to(DEV).log_prob(xc, context=c)
        g = torch.autograd.grad(lp.mean(), xc)[0]
        assert lp.shape == batch_shape and g.shape == x.shape and g.isfinite().all()


@pytest.mark.parametrize('name', ['ElementwiseAffine', 'ElementwiseShift', 'ElementwiseRQSpline', 'ActNorm'])
@pytest.mark.parametrize('context_shape', CONTEXT_SHAPES)
def test_elementwise_with_context(name, context_shape):
    from torchflows_b200.bijections.finite.autoregressive import layers
    torch.manual_seed(0)
    event_shape, batch_shape = (3, 5, 2), (5, 2)
    layer = getattr(layers, name)(event_shape, context_shape=context_shape).to(DEV)
    x = torch.randn(*batch_shape, *event_shape, device=DEV)
    c = torch.randn(*batch_shape, *context_shape, device=DEV)
    _z, _ = layer.forward(x, context=c)
    xr, _ = layer.inverse(_z, context=c)
    assert torch.allclose(x, xr, atol=ATOL)


def test_fit_and_sample_with_context():
    """test/test_fit.py:156-182 and flows.py:680-692 (both context broadcasting options of Flow.sample)."""
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import RealNVP
    torch.manual_seed(0)
    for n_train, context_shape in ((10, (2,)), (2200, (3, 5, 2))):
        flow = Flow(RealNVP((3,), context_shape=context_shape)).to(DEV)
        x, c = torch.randn(n_train, 3), torch.randn(n_train, *context_shape)
        flow.fit(x, n_epochs=2, context_train=c, x_val=torch.randn(5, 3), context_val=torch.randn(5, *context_shape))
        a = flow.sample(4, context=torch.randn(4, *context_shape))               # one context per sample
        b = flow.sample(6, context=torch.randn(2, *context_shape))               # 6 samples for each of 2 contexts
        assert a.shape == (4, 3) and b.shape == (6, 2, 3)
