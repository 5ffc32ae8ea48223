"""GPU tests of the backward kernel and Flow.fit against reference autograd / the reference's fit loop."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def build(preset, event_shape, kwargs, state_dict, dev):
    import torchflows_b200.architectures as arch
    from torchflows_b200 import Flow
    flow = Flow(getattr(arch, preset)(event_shape, **kwargs))
    flow.load_state_dict(state_dict)
    return flow.to(dev)


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize('idx', range(7))
def test_gradients_vs_reference_autograd(golden, dev, idx):
    """P4: loss = _base_batch_loss (flows.py:199-224); every parameter gradient and dloss/dx against the
    reference's autograd (fp32).  Bar: relative L2 per tensor 1e-3 (the reference's own fp32 gradients are 6e-4
    from its fp64 gradients for spline knots, tests/test_c_oracle_and_hostmath.py), 1e-4 for affine flows."""
    c = golden('grads.pt')[idx]
    flow = build(c['preset'], c['event_shape'], c['kwargs'], c['state_dict'], dev).eval()
    x = c['x'].to(dev).requires_grad_(True)
    loss = flow._base_batch_loss((x, c['w'].to(dev)))
    loss.backward()
    assert abs(float(loss.detach()) - float(c['loss'])) <= 1e-5 * (1 + abs(float(c['loss'])))
    tol = 2e-3 if 'RQNSF' in c['preset'] else 1e-4
    assert rel(x.grad, c['grad_x']) < tol, ('grad_x', rel(x.grad, c['grad_x']))
    params = dict(flow.named_parameters())
    checked = 0
    for k, g in c['grads'].items():
        if g.numel() == 0:
            continue
        assert params[k].grad is not None, k
        if g.norm() == 0:
            assert params[k].grad.abs().max().item() < 1e-6, k
        else:
            assert rel(params[k].grad, g) < tol, (k, rel(params[k].grad, g))
        checked += 1
    assert checked >= 6


def test_input_gradient_is_finite_like_reference_test(dev):
    """test/test_autograd_bijections.py:41-54 re-pointed at this package."""
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import NICE, RealNVP, CouplingRQNSF, MAF, MaskedAutoregressiveRQNSF
    for cls in (NICE, RealNVP, CouplingRQNSF, MAF, MaskedAutoregressiveRQNSF):
        for batch_shape, event_shape in (((1,), (2,)), ((5, 2, 3), (3, 5, 2))):
            torch.manual_seed(0)
            flow = Flow(cls(event_shape)).to(dev)
            x = torch.randn(*batch_shape, *event_shape, device=dev, requires_grad=True)
            lp = flow.log_prob(x)
            assert lp.shape == batch_shape and torch.isfinite(lp).all()
            g = torch.autograd.grad(lp.mean(), x)[0]
            assert g.shape == x.shape and torch.isfinite(g).all()


@pytest.mark.parametrize('idx', range(4))
def test_fit_loss_trajectory(golden, dev, idx):
    """P5: the inner loop of BaseFlow.fit (flows.py:379-398), full batch, 20 AdamW steps from identical weights.
    Cases 0-2 keep the constructor's ActNorm parameters: loss trajectory within 1e-3 relative of the reference's.
    Case 3 lets ActNorm data-initialise on the first step; several gradients are then rounding noise that Adam's
    normalisation amplifies (see tests/golden/make_golden.py), so only the first loss is compared tightly and the
    rest of the trajectory loosely."""
    c = golden('fit.pt')[idx]
    flow = build(c['preset'], c['event_shape'], {}, c['state_dict0'], dev)
    flow.train()
    if not c['data_init']:
        for layer in flow.bijection.layers:
            if hasattr(layer, 'first_training_batch_pass'):
                layer.first_training_batch_pass = False
    x = c['x'].to(dev)
    w = torch.ones(len(x), device=dev)
    opt = torch.optim.AdamW(flow.parameters(), lr=c['lr'])
    losses = []
    for _ in range(20):
        opt.zero_grad()
        loss = flow._base_batch_loss((x, w))
        losses.append(float(loss.detach()))
        loss.backward()
        opt.step()
    for i, (a, b) in enumerate(zip(losses, c['losses'])):
        tol = 1e-3 if (not c['data_init'] or i == 0) else 3e-2
        assert abs(a - b) <= tol * (1 + abs(b)), (i, a, b)
    if not c['data_init']:
        flow.eval()
        with torch.no_grad():
            lp = flow.log_prob(x).cpu()
        assert (lp - c['log_prob20']).abs().max().item() <= 2e-2 * (1 + c['log_prob20'].abs().max().item())


def test_readme_example(dev):
    """README.md:9-31: fit RealNVP(3) to 1000 standard-normal points, then log_prob and sample(50)."""
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import RealNVP
    torch.manual_seed(0)
    n_data, n_dim = 1000, 3
    x = torch.randn(n_data, n_dim)
    flow = Flow(RealNVP(n_dim)).to(dev)
    flow.fit(x, show_progress=False)
    with torch.no_grad():
        log_prob = flow.log_prob(x)
        x_new = flow.sample(50)
    assert log_prob.shape == (n_data,) and x_new.shape == (50, n_dim)
    assert abs(float(log_prob.mean()) - (-4.26)) < 0.1       # reference: -4.263, true N(0,I): -4.265
    assert not flow.training


def test_fit_gaussian_statistical(dev):
    """test/test_fit.py:81-126 (local_only in the reference), shortened: after fitting N(0, diag sigma^2) the
    sample std is within rtol 0.1 of sigma."""
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import RealNVP, CouplingRQNSF
    for cls in (RealNVP, CouplingRQNSF):
        torch.manual_seed(0)
        sigma = torch.tensor([0.1, 1.0, 10.0])
        x = torch.randn(10000, 3) * sigma
        flow = Flow(cls(3)).to(dev)
        flow.fit(x, n_epochs=150, lr=0.05 if cls is RealNVP else 0.01)
        with torch.no_grad():
            s = flow.sample(100000).std(dim=0).cpu()
        assert torch.allclose(s, sigma, rtol=0.1), (cls.__name__, s)


def test_fit_options_smoke(dev):
    """test/test_fit.py:132-182 shapes of the API: tiny data, validation data, early stopping, adaptive batches,
    deepcopy before / after (test/test_deepcopy.py)."""
    from copy import deepcopy
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import MAF, NICE
    torch.manual_seed(0)
    for n_train in (1, 10, 2200):
        flow = Flow(NICE((2, 3))).to(dev)
        deepcopy(flow)
        flow.fit(torch.randn(n_train, 2, 3), n_epochs=2, x_val=torch.randn(7, 2, 3), early_stopping=True)
        deepcopy(flow)
    flow = Flow(MAF(4)).to(dev)
    flow.fit(torch.randn(5000, 4), n_epochs=12, batch_size='adaptive', w_train=torch.rand(5000))
    assert not flow.training


def test_composite_path_matches_fused_path(dev):
    """Layers outside the fused kernels' budget (e.g. D=1024 with n_hidden=1024) run as a composite: conditioner as
    library GEMMs + stand-alone transformer kernels.  Same weights through both paths must agree (values and gradients)."""
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import CouplingRQNSF, RealNVP
    for cls, D in ((CouplingRQNSF, 16), (RealNVP, 10)):
        torch.manual_seed(3)
        fused = Flow(cls(D)).to(dev).eval()
        comp = Flow(cls(D)).to(dev).eval()
        comp.load_state_dict(fused.state_dict())
        for layer in comp.bijection.layers:
            if hasattr(layer, '_fusable'):
                layer._fusable = False
            if hasattr(layer, 'use_global_parameters'):
                layer.lower = lambda direction: None          # elementwise layers through the transformer kernel
        assert comp.bijection.lower('forward') is None
        x = torch.randn(257, D, device=dev)
        outs = []
        for flow in (fused, comp):
            xi = x.clone().requires_grad_(True)
            loss = flow._base_batch_loss((xi, torch.ones(257, device=dev)))
            loss.backward()
            outs.append((loss.detach(), xi.grad, {k: p.grad for k, p in flow.named_parameters() if p.grad is not None}))
            with torch.no_grad():
                outs[-1] += (flow._sample_from_base(x, no_grad=True),)
        assert abs(float(outs[0][0]) - float(outs[1][0])) < 1e-5 * (1 + abs(float(outs[0][0])))
        assert rel(outs[1][1], outs[0][1]) < 1e-3
        for k, g in outs[0][2].items():
            if g.numel() and g.norm() > 0:
                assert rel(outs[1][2][k], g) < 2e-3, k
        assert rel(outs[1][3], outs[0][3]) < 1e-4


def test_wide_config_fit_step_runs(dev):
    """BASELINE configs[4] shape (scaled-down batch): one fit step of CouplingRQNSF(1024, n_hidden=1024) is finite."""
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import CouplingRQNSF
    torch.manual_seed(0)
    flow = Flow(CouplingRQNSF(1024, conditioner_kwargs={'n_hidden': 1024})).to(dev)
    assert sum(p.numel() for p in flow.parameters() if p.requires_grad) == 25195520
    x = torch.randn(512, 1024, device=dev)
    flow.train()
    l0 = float(flow.train_step(x))
    l1 = float(flow.train_step(x))
    assert l0 == l0 and l1 == l1 and abs(l1) < 1e6


def test_sampling_direction_gradients_vs_oracle_autograd(dev):
    """gradients through Flow.sample (inverse direction, incl. the inverse spline): loss = mean(sum(x^2)) +
    mean(sample log_prob) against torch autograd through the CPU oracle on the same base noise."""
    from oracle.flow_oracle import OracleFlow
    from torchflows_b200 import Flow
    import torchflows_b200.architectures as arch
    for preset, D in (('RealNVP', 8), ('NICE', 6), ('CouplingRQNSF', 8), ('IAF', 6), ('CouplingRQNSF', 32),
                      ('CouplingRQNSF', 3), ('RealNVP', 3), ('CouplingRQNSF', 5), ('MAF', 6), ('MAF', 33)):
        torch.manual_seed(5)
        flow = Flow(getattr(arch, preset)(D)).eval()
        sd = {k: v.clone().requires_grad_(v.is_floating_point() and ('weight' in k or 'bias' in k or 'value' in k))
              for k, v in flow.state_dict().items()}
        o = OracleFlow(preset, (D,), {})
        o.sd = sd
        noise = torch.randn(64, D)
        xo, lpo = o.sample_from_noise(noise, return_log_prob=True)
        (xo.pow(2).sum(-1).mean() + lpo.mean()).backward()
        flow = flow.to(dev)
        x, lp = flow._sample_from_base(noise.to(dev), return_log_prob=True)
        loss = x.pow(2).sum(-1).mean() + lp.mean()
        loss.backward()
        tol = 5e-3 if 'RQNSF' in preset else 2e-4
        checked = 0
        for k, p in flow.named_parameters():
            g_ref = sd[k].grad
            if p.grad is None or g_ref is None or g_ref.norm() == 0:
                continue
            assert rel(p.grad, g_ref) < tol, (preset, k, rel(p.grad, g_ref))
            checked += 1
        assert checked >= 6, preset
        assert rel(x, xo) < 1e-4 and rel(lp, lpo) < 1e-4, preset


def test_variational_fit_and_kl_fit(dev):
    """flows.py:96-197, 495-603: SVI towards N(mu, sigma^2 I) recovers mean and scale; KL(p||q) fit runs."""
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import RealNVP, CouplingRQNSF, MAF
    torch.manual_seed(0)
    mu, sigma = torch.tensor([1.0, -2.0, 0.5], device=dev), 0.5

    def target_log_prob(x):
        return -0.5 * (((x - mu) / sigma) ** 2).sum(-1)
    def svi_recovers_target(cls, seed, n_epochs):
        torch.manual_seed(seed)
        flow = Flow(cls(3)).to(dev)
        flow.variational_fit(target_log_prob, n_epochs=n_epochs, lr=0.05, n_samples=256, check_for_divergences=True)
        assert not flow.training
        with torch.no_grad():
            s = flow.sample(20000)
        # a stochastic optimiser from a random initialisation: generous bounds (the reference reaches ~0.03 / ~0.15)
        return (s.mean(0) - mu).abs().max().item() < 0.25 and (s.std(0) - sigma).abs().max().item() < 0.25

    # lr = 0.05 from a random initialisation occasionally parks a coordinate in a poor optimum (the reference does too, and
    # which seeds do depends on last-bit details of the optimiser): the claim tested is that SVI works, so one of three
    # initialisations has to recover the target
    for cls in (RealNVP, CouplingRQNSF):
        assert any(svi_recovers_target(cls, seed, 600) for seed in (1, 2, 3)), cls.__name__
    flow = Flow(MAF(3)).to(dev)
    x = torch.randn(2000, 3) * sigma + mu.cpu()
    flow.fit_kl_p_to_q(x[:1500], x[1500:], lambda t: -target_log_prob(t.to(dev)).cpu(), n_epochs=5, lr=0.01)
    # MAF samples through the sequential direction: fused backward as well
    assert any(svi_recovers_target(MAF, seed, 400) for seed in (2, 3, 4))


@pytest.mark.parametrize('preset,D,n,bs', [('RealNVP', 3, 500, None), ('CouplingRQNSF', 16, 1024, 256), ('MAF', 8, 700, 256)])
def test_fit_with_cuda_graph_matches_eager_fit(preset, D, n, bs):
    """fit(cuda_graph=True): the training step replayed from a CUDA graph (full batches) mixed with eager steps (the first
    three, and the ragged last batch of every epoch) follows the eager trajectory; validation and the best-weights
    snapshot see the replayed updates."""
    from torchflows_b200 import Flow
    import torchflows_b200.architectures as arch
    dev = torch.device('cuda:0')
    finals = []
    for graph in (False, True):
        torch.manual_seed(0)
        x = torch.randn(n, D) * 1.5 + 0.3
        torch.manual_seed(1)
        flow = Flow(getattr(arch, preset)(D)).to(dev)
        flow.fit(x, n_epochs=30, batch_size=bs, lr=0.01, x_val=x[:100], cuda_graph=graph)
        assert not flow.training
        with torch.no_grad():
            finals.append(flow.log_prob(x.to(dev)).mean().item())
    assert abs(finals[0] - finals[1]) < 5e-3 * (1 + abs(finals[0])), finals


def test_flow_apply_saving_reports_whether_layer_inputs_were_saved():
    """b2f_flow_apply_saving: the tensor-core kernel writes the conditioner-layer inputs (first layer's = the flow input
    after the leading elementwise run; checked for an identity ElementwiseAffine), the other kernels report saved = 0."""
    from torchflows_b200 import Flow, _native as N, _program as P
    from torchflows_b200.architectures import CouplingRQNSF, RealNVP
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    x = torch.randn(300, 64, device=dev)
    for cls, expect in ((CouplingRQNSF, True), (RealNVP, False)):
        flow = Flow(cls(64)).to(dev).eval()
        ops = flow.bijection.fused_ops('forward')
        with torch.no_grad():
            y, ld, lp, ws = N.flow_apply(P.op_dicts(ops, D=64), x, True, True, False, None, None, 0, save_layer_inputs=True)
            y2, ld2, _ = N.flow_apply(P.op_dicts(ops, D=64), x, True, True, False, None, None, 0)
        assert (ws is not None) == expect, cls.__name__
        assert torch.equal(y, y2) and torch.equal(ld, ld2)
        if ws is not None:
            n_cond = sum(1 for o in ops if o.kind in (N.OP_COUPLING, N.OP_MADE, N.OP_MADE_SEQ))
            saved = ws[:n_cond * 300 * 64].view(n_cond, 300, 64)
            # what enters the first coupling layer is the leading ElementwiseAffine applied to x (the flip in between only
            # relabels columns)
            assert ops[0].kind == N.OP_ELEMENTWISE and ops[1].kind == N.OP_FLIP
            with torch.no_grad():
                first_in = P.run_program([ops[0]], x)[0]
            assert torch.allclose(saved[0], first_in, atol=1e-5)
            assert torch.isfinite(saved).all()


# ---------------------------------------------------------------------------------------------------------
# P4 at the benchmark shapes, with the fp64 oracle as referee (SURVEY 8c iii)
# ---------------------------------------------------------------------------------------------------------
def _oracle_grads(preset, D, sd, x, dtype, kwargs_hidden=None, chunk=512):
    """Gradients of _base_batch_loss (flows.py:199-224) by torch autograd through the CPU oracle, accumulated over
    chunks of the batch (the loss is a mean over rows plus a batch-independent regulariser)."""
    from oracle.flow_oracle import OracleFlow
    leaves = {k: v.detach().clone().to(dtype).requires_grad_(True) if v.is_floating_point() and v.numel() > 0 else v.clone()
              for k, v in sd.items()}
    o = OracleFlow(preset, (D,), {})
    o.sd = leaves
    xg = x.detach().clone().to(dtype).requires_grad_(True)
    B = x.shape[0]
    total = 0.0
    for s in range(0, B, chunk):
        part = -(o.log_prob(xg[s:s + chunk]).sum()) / B / D
        part.backward()
        total += float(part)
    reg = o.regularization()
    reg.backward()
    grads = {k: v.grad for k, v in leaves.items() if isinstance(v, torch.Tensor) and v.requires_grad and v.grad is not None}
    return total + float(reg), xg.grad, grads


@pytest.mark.parametrize('name,D,kwargs', [('Q256', 256, {}), ('W1024', 1024, {'conditioner_kwargs': {'n_hidden': 1024}})])
def test_gradients_at_benchmark_shapes_with_fp64_referee(dev, name, D, kwargs):
    """CouplingRQNSF(256) (fused backward kernel) and CouplingRQNSF(1024, n_hidden=1024) (BASELINE configs[4]) at
    B = 4096: loss, dloss/dx and every parameter gradient.  Where our gradient is further than 1e-4 (relative L2) from the
    reference's fp32 autograd, the fp64 oracle referees: we must be as close to fp64 as 3x the fp32 reference is itself
    (its spline-knot gradients are ~6e-4 from fp64, tests/test_c_oracle_and_hostmath.py)."""
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import CouplingRQNSF
    torch.manual_seed(7)
    flow = Flow(CouplingRQNSF(D, **kwargs)).eval()
    sd = {k: v.detach().clone() for k, v in flow.state_dict().items()}
    g = torch.Generator().manual_seed(8)
    B = 4096
    x = torch.randn(B, D, generator=g) * 1.2
    loss32, gx32, g32 = _oracle_grads('CouplingRQNSF', D, sd, x, torch.float32)
    loss64, gx64, g64 = _oracle_grads('CouplingRQNSF', D, sd, x, torch.float64)
    flow = flow.to(dev)
    xd = x.to(dev).requires_grad_(True)
    loss = flow._base_batch_loss((xd, torch.ones(B, device=dev)))
    loss.backward()
    assert abs(float(loss.detach()) - loss64) <= 1e-4 * (1 + abs(loss64)), (float(loss.detach()), loss64)

    def check(what, ours, r32, r64):
        if r64.norm() == 0:
            assert ours.abs().max().item() < 1e-6, what
            return
        e32 = rel(ours, r32)
        if e32 < 1e-4:
            return
        e64, ref_e64 = rel(ours, r64), rel(r32, r64)
        assert e64 <= max(1e-4, 3.0 * ref_e64), f'{name} {what}: ours vs fp32 {e32:.2e}, ours vs fp64 {e64:.2e}, fp32 vs fp64 {ref_e64:.2e}'

    check('grad_x', xd.grad, gx32, gx64)
    checked = 0
    for k, p in flow.named_parameters():
        if p.grad is None or k not in g64:
            continue
        check(k, p.grad, g32[k], g64[k])
        checked += 1
    assert checked >= 10


def test_fit_steps_on_layers_beyond_the_fused_backward_budget(dev):
    """RQ coupling at D = 512 and MAF at D = 1024 must TRAIN (the backward kernel's shared-memory
    footprint decides fusability, with smaller tile shapes and a composite path as fallbacks), and
    InverseAutoregressiveRQNSF must fit with default settings (exact log-det when gradients are needed)."""
    import warnings
    from torchflows_b200 import Flow
    from torchflows_b200.architectures import CouplingRQNSF, MAF, InverseAutoregressiveRQNSF
    for cls, D, n in ((CouplingRQNSF, 512, 256), (CouplingRQNSF, 448, 256), (MAF, 1024, 256), (MAF, 768, 256),
                      (InverseAutoregressiveRQNSF, 6, 200)):
        torch.manual_seed(0)
        flow = Flow(cls(D)).to(dev)
        x = torch.randn(n, D, device=dev)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            flow.fit(x, n_epochs=3, lr=1e-3, batch_size=None)
        with torch.no_grad():
            lp = flow.log_prob(x)
        assert torch.isfinite(lp).all(), (cls.__name__, D)
        for p in flow.parameters():
            assert torch.isfinite(p).all()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_data_parallel_fit_over_nccl_matches_single_process():
    """2 ranks over NCCL, per-step loss equal to the single-process run within 1e-4 relative for 20 steps,
    weights broadcast from rank 0 (ranks construct their flows under different seeds), identical weights at the end."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', str(29600 + os.getpid() % 300), os.path.join(root, 'scripts', 'dp_fit_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and 'DP_FIT_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
