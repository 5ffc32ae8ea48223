"""Flow.sample with the library's own base draws (include/b2f.h b2f_flow_sample / b2f_philox_normal,
csrc/b2f_philox.cuh): the counter-based stream against a numpy restatement of Philox4x32-10 + Box-Muller (known-answer
vector of the Random123 distribution included), determinism, the fused in-kernel draw against the materialised stream
pushed through the given-noise path, and moments of the draws.
Reference behaviour: base_distributions/gaussian.py:41-44 (torch.randn on the CPU) + flows.py:660-713."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def philox4x32_10(counter, seed):
    """numpy restatement (uint64 arithmetic) of csrc/b2f_philox.cuh; counter: array of 64-bit counters."""
    c = [np.asarray(counter, dtype=np.uint64) & np.uint64(0xFFFFFFFF), np.asarray(counter, dtype=np.uint64) >> np.uint64(32),
         np.zeros_like(counter, dtype=np.uint64), np.zeros_like(counter, dtype=np.uint64)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c[0], np.uint64(0xCD9E8D57) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & m32, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & m32]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & m32, (k1 + np.uint64(0xBB67AE85)) & m32
    return c


def normals(counter, seed):
    u = philox4x32_10(counter, seed)
    out = []
    for a, b in ((u[0], u[1]), (u[2], u[3])):
        u1 = ((a >> np.uint64(8)).astype(np.float64) + 0.5) * 2.0 ** -24
        th = (b >> np.uint64(8)).astype(np.float64) * (2 * np.pi * 2.0 ** -24)
        r = np.sqrt(-2 * np.log(u1))
        out += [r * np.cos(th), r * np.sin(th)]
    return np.stack(out, axis=-1)        # (..., 4)


def test_numpy_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10, counter 0, key 0
    u = philox4x32_10(np.zeros(1, dtype=np.uint64), 0)
    assert [int(v[0]) for v in u] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


@pytest.mark.parametrize('n_rows,D,seed,offset', [(64, 8, 1234, 0), (33, 6, 2 ** 40 + 17, 5), (1000, 256, 7, 2 ** 33)])
def test_philox_normal_matches_the_restatement(n_rows, D, seed, offset):
    from torchflows_b200 import _native as N
    dev = torch.device('cuda:0')
    z = N.philox_normal(n_rows, D, dev, seed, offset).cpu().double().numpy().reshape(-1)
    n4 = (n_rows * D + 3) // 4
    ref = normals(np.arange(n4, dtype=np.uint64) + np.uint64(offset), seed).reshape(-1)[:n_rows * D]
    assert np.abs(z - ref).max() < 2e-5            # __logf / __sincosf of the device versus float64
    loc, ls = torch.randn(D, device=dev), torch.randn(D, device=dev) * 0.3
    z2 = N.philox_normal(n_rows, D, dev, seed, offset, loc, ls).cpu().double().numpy()
    ref2 = loc.cpu().double().numpy() + np.exp(ls.cpu().double().numpy()) * ref.reshape(n_rows, D)
    assert np.abs(z2 - ref2).max() < 1e-4


def test_philox_normal_moments():
    from torchflows_b200 import _native as N
    z = N.philox_normal(1 << 16, 256, torch.device('cuda:0'), 99, 0).double()
    n = z.numel()
    assert abs(float(z.mean())) < 5 / n ** 0.5
    assert abs(float(z.var()) - 1) < 5 * (2 / n) ** 0.5
    assert abs(float((z ** 3).mean())) < 5 * (15 / n) ** 0.5
    assert abs(float((z ** 4).mean()) - 3) < 5 * (96 / n) ** 0.5
    # Kolmogorov-Smirnov against the normal CDF on a subsample
    s = torch.sort(z.reshape(-1)[:200000]).values
    cdf = 0.5 * (1 + torch.erf(s / 2 ** 0.5))
    emp = torch.arange(1, s.numel() + 1, device=s.device, dtype=torch.float64) / s.numel()
    assert float((cdf - emp).abs().max()) < 1.95 / s.numel() ** 0.5      # alpha ~ 0.001
    # columns are independent draws, rows too
    c = torch.corrcoef(z[:, :8].T)
    assert float((c - torch.eye(8, device=c.device, dtype=c.dtype)).abs().max()) < 0.03


@pytest.mark.parametrize('preset,D,n', [('CouplingRQNSF', 64, 1000), ('CouplingRQNSF', 256, 20000), ('RealNVP', 16, 777),
                                        ('MAF', 8, 300), ('MaskedAutoregressiveRQNSF', 12, 200)])
def test_sample_is_deterministic_and_equals_given_noise_path(preset, D, n):
    """Same torch seed -> same samples; and the fused draw (spline coupling programs: noise made in registers by the kernel)
    equals the materialised stream pushed through the given-noise path (which is golden-checked against the reference)."""
    from torchflows_b200 import Flow, _native as N, architectures as A
    from torchflows_b200 import _program as prog
    dev = torch.device('cuda:0')
    torch.manual_seed(3)
    flow = Flow(getattr(A, preset)(D)).to(dev).eval()
    with torch.no_grad():
        torch.manual_seed(11)
        x1, lp1 = flow.sample((n,), no_grad=True, return_log_prob=True)
        kernel = N.last_flow_kernel()
        x1b = flow.sample((n,), no_grad=True)            # the stream advances: different draws
        torch.manual_seed(11)
        x2, lp2 = flow.sample((n,), no_grad=True, return_log_prob=True)
        assert torch.equal(x1, x2) and torch.equal(lp1, lp2)
        assert not torch.equal(x1, x1b)
        if preset == 'CouplingRQNSF':
            assert kernel == N.KERNEL_TCQ
        torch.manual_seed(11)
        seed, offset = prog.next_noise_stream()
        z = N.philox_normal(n, D, dev, seed, offset)
        x3, lp3 = flow._sample_from_base(z, no_grad=True, return_log_prob=True)
        assert float((x3 - x1).abs().max()) <= 1e-5 * (1 + float(x1.abs().max()))
        assert float((lp3 - lp1).abs().max()) <= 1e-4 * (1 + float(lp1.abs().max()))


def test_sample_with_a_non_standard_base():
    from torchflows_b200 import Flow, _native as N, architectures as A
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    flow = Flow(A.CouplingRQNSF(32)).to(dev).eval()
    with torch.no_grad():
        flow.base.loc.copy_(torch.randn(32, device=dev))
        flow.base.log_scale.copy_(torch.randn(32, device=dev) * 0.2)
        torch.manual_seed(5)
        x, lp = flow.sample((4096,), no_grad=True, return_log_prob=True)
        torch.manual_seed(5)
        from torchflows_b200 import _program as prog
        seed, offset = prog.next_noise_stream()
        z = N.philox_normal(4096, 32, dev, seed, offset, flow.base.loc.reshape(-1), flow.base.log_scale.reshape(-1))
        x3, lp3 = flow._sample_from_base(z, no_grad=True, return_log_prob=True)
        assert float((x3 - x).abs().max()) <= 1e-5 * (1 + float(x.abs().max()))
        assert float((lp3 - lp).abs().max()) <= 1e-4 * (1 + float(lp.abs().max()))
        # (lp is base_log_prob(z) + log_det_inverse, the reference's flows.py:710-712 -- deliberately not compared with
        # flow.log_prob(x), which subtracts that log-determinant)
