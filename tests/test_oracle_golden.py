"""Pin oracle/flow_oracle.py against vectors produced by the real reference (tests/golden/make_golden.py)."""
import pytest
import torch

from oracle import flow_oracle as fo


def _eq(a, b, what):
    assert a.shape == b.shape, what
    assert torch.equal(a, b), f'{what}: max abs diff {(a - b).abs().max().item():.3e}'


def test_transformers_bit_identical(golden):
    torch.set_num_threads(1)
    for c in golden('transformers.pt'):
        ne = len(c['event_shape'])
        kw = dict(n_bins=c['n_bins'], boundary=c['boundary']) if c['kind'] == 'rq' else {}
        if c['kind'] == 'rq':
            z, ld, k = fo.rq_forward(c['x'], c['h'], n_event_dims=ne, return_bins=True, **kw)
            xr, ldi, ki = fo.rq_inverse(c['z'], c['h'], n_event_dims=ne, return_bins=True, **kw)
            _eq(k, c['k_f'], 'k_f')
            _eq(ki, c['k_i'], 'k_i')
        else:
            fwd, inv = fo.TRANSFORMERS[c['kind']]
            z, ld = fwd(c['x'], c['h'], ne)
            xr, ldi = inv(c['z'], c['h'], ne)
        _eq(z, c['z'], 'z')
        _eq(ld, c['ld_f'], 'ld_f')
        _eq(xr, c['xr'], 'xr')
        _eq(ldi, c['ld_i'], 'ld_i')


@pytest.mark.parametrize('idx', range(16))
def test_presets_bit_identical(golden, idx):
    torch.set_num_threads(1)
    c = golden('presets.pt')[idx]
    o = fo.OracleFlow(c['preset'], c['event_shape'], c['state_dict'])
    z, ld = o.forward(c['x'])
    _eq(z, c['z'], 'z')
    _eq(ld, c['ld_f'], 'ld_f')
    _eq(o.log_prob(c['x']), c['log_prob'], 'log_prob')
    xs, lps = o.sample_from_noise(c['noise'], return_log_prob=True)
    _eq(xs, c['xs'], 'xs')
    _eq(lps, c['lp_s'], 'lp_s')
    xr, ldr = o.inverse(c['z'])
    _eq(xr, c['xr'], 'xr')
    _eq(ldr, c['ld_r'], 'ld_r')
    # state T: ActNorm data-dependent initialisation
    o.actnorm_initialise(c['x'])
    for k, v in c['actnorm_T'].items():
        _eq(o.sd[k], v, k)
    _eq(o.log_prob(c['x']), c['log_prob_T'], 'log_prob_T')


def test_gradients_match_reference_autograd(golden):
    torch.set_num_threads(1)
    for c in golden('grads.pt'):
        sd = {k: v.clone().requires_grad_(v.is_floating_point() and ('weight' in k or 'bias' in k or 'value' in k))
              for k, v in c['state_dict'].items()}
        o = fo.OracleFlow(c['preset'], c['event_shape'], {})
        o.sd = sd
        x = c['x'].clone().requires_grad_(True)
        loss = o.batch_loss(x, c['w'])
        loss.backward()
        assert torch.allclose(loss.detach(), c['loss'], rtol=1e-6, atol=1e-6)
        assert torch.allclose(x.grad, c['grad_x'], rtol=1e-5, atol=1e-7)
        for k, g in c['grads'].items():
            if g.numel() == 0:
                continue
            assert torch.allclose(sd[k].grad, g, rtol=1e-5, atol=1e-7), k


def test_lrs_and_scale_transformers_bit_identical(golden):
    """Linear rational spline / Scale restatements against the reference's outputs (tests/golden/make_golden_lrs.py)."""
    torch.set_num_threads(1)
    for c in golden('lrs.pt')['transformers']:
        ne = len(c['event_shape'])
        if c['kind'] == 'lrs':
            o, l = fo.lrs_forward(c['x'], c['h'], c['n_bins'], c['boundary'], ne)
            o2, l2 = fo.lrs_inverse(c['z_in'], c['h'], c['n_bins'], c['boundary'], ne)
        else:
            o, l = fo.scale_forward(c['x'], c['h'], ne)
            o2, l2 = fo.scale_inverse(c['z_in'], c['h'], ne)
        _eq(o, c['forward']['out'], 'forward out')
        _eq(l, c['forward']['ld'], 'forward log_det')
        _eq(o2, c['inverse']['out'], 'inverse out')
        _eq(l2, c['inverse']['ld'], 'inverse log_det')


def test_lrs_presets_bit_identical(golden):
    torch.set_num_threads(1)
    for c in golden('lrs.pt')['presets']:
        o = fo.OracleFlow(c['preset'], c['event_shape'], c['state_dict'])
        _eq(o.log_prob(c['x']), c['log_prob'], c['preset'] + ' log_prob')
        xs, lps = o.sample_from_noise(c['noise'], return_log_prob=True)
        _eq(xs, c['xs'], c['preset'] + ' xs')
        _eq(lps, c['lp_s'], c['preset'] + ' lp_s')
