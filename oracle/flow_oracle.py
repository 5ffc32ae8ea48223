"""CPU oracle for the coupling / masked-autoregressive hot path of torchflows v1.2.0.

TEST INFRASTRUCTURE ONLY.  Nothing under ``torchflows_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` do, and there only as the checker / CPU baseline, never as the thing measured or shipped.

What it is: a functional restatement (plain functions over a ``state_dict`` of tensors, no
``nn.Module`` tree) of the reference algorithm, written with the same ATen CPU ops in the same
order as the reference, because the reference's arithmetic *is* PyTorch eager on CPU
(SURVEY.md section 8c).  It is therefore expected to be bit-identical to the reference on CPU.

Pinning: ``tests/golden/make_golden.py`` imports the real reference from ``/root/reference`` (in the
build container) and stores inputs, weights and outputs under ``tests/golden/*.pt``;
``tests/test_oracle_golden.py`` checks this file against those vectors (exact equality for the
one-pass paths).  Parity is therefore pinned by outputs of the reference itself.

Citations are ``file:line`` relative to ``/root/reference/torchflows``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------------
# Elementwise transformers
# ----------------------------------------------------------------------------------------------
AFFINE_MIN_SCALE = 1e-10  # bijections/finite/autoregressive/transformers/linear/affine.py:19


def affine_scale(u_alpha: Tensor, m: float = AFFINE_MIN_SCALE) -> Tensor:
    """alpha = exp(log(1-m) + u/2) + m   (affine.py:33-34)."""
    return torch.exp(math.log(1 - m) + u_alpha / 2) + m


def affine_unconstrain_scale(scale: Tensor, m: float = AFFINE_MIN_SCALE) -> Tensor:
    """affine.py:36-37."""
    return (torch.log(scale - m) - math.log(1 - m)) * 2


def _sum_event(t: Tensor, n_event_dims: int) -> Tensor:
    """utils.py:158 (sum_except_batch)."""
    return torch.sum(t, dim=list(range(t.dim()))[-n_event_dims:])


def affine_forward(x: Tensor, h: Tensor, n_event_dims: int = 1) -> Tuple[Tensor, Tensor]:
    """z = alpha*x + beta, log_det = sum log alpha   (affine.py:39-48)."""
    alpha = affine_scale(h[..., 0])
    log_alpha = torch.log(alpha)
    return alpha * x + h[..., 1], _sum_event(log_alpha, n_event_dims)


def affine_inverse(z: Tensor, h: Tensor, n_event_dims: int = 1) -> Tuple[Tensor, Tensor]:
    """x = (z-beta)/alpha, log_det = -sum log alpha   (affine.py:50-59)."""
    alpha = affine_scale(h[..., 0])
    log_alpha = torch.log(alpha)
    return (z - h[..., 1]) / alpha, -_sum_event(log_alpha, n_event_dims)


def shift_forward(x: Tensor, h: Tensor, n_event_dims: int = 1) -> Tuple[Tensor, Tensor]:
    """affine.py:149-153."""
    return x + h[..., 0], torch.zeros(x.shape[: x.dim() - n_event_dims], device=x.device)


def shift_inverse(z: Tensor, h: Tensor, n_event_dims: int = 1) -> Tuple[Tensor, Tensor]:
    """affine.py:155-159."""
    return z - h[..., 0], torch.zeros(z.shape[: z.dim() - n_event_dims], device=z.device)


# ----------------------------------------------------------------------------------------------
# Rational-quadratic spline   (transformers/spline/rational_quadratic.py, spline/base.py)
# ----------------------------------------------------------------------------------------------
RQ_MIN_BIN = 1e-3      # rational_quadratic.py:36
RQ_MIN_DELTA = 1e-5    # rational_quadratic.py:37
RQ_EDGE_U = math.log(math.expm1(1 - RQ_MIN_DELTA))  # rational_quadratic.py:38


def rq_bins(u: Tensor, lo: float, hi: float) -> Tuple[Tensor, Tensor]:
    """Knot positions and bin sizes from logits (rational_quadratic.py:45-54)."""
    n_bins = u.shape[-1]
    sizes = torch.softmax(u, dim=-1)
    sizes = RQ_MIN_BIN + (1 - RQ_MIN_BIN * n_bins) * sizes
    knots = torch.cumsum(sizes, dim=-1)
    knots = F.pad(knots, pad=(1, 0), mode='constant', value=0.0)
    knots = (hi - lo) * knots + lo
    knots[..., 0] = lo
    knots[..., -1] = hi
    sizes = knots[..., 1:] - knots[..., :-1]
    return knots, sizes


def _rq_log_det(s, d0, d1, xi, q, t1):
    """rational_quadratic.py:56-63."""
    log_num = 2 * torch.log(s) + torch.log(d1 * xi ** 2 + 2 * s * q + d0 * (1 - xi) ** 2)
    log_den = 2 * torch.log(s + t1 * q)
    return log_num - log_den


def _rq_split(h: Tensor, n_bins: int):
    """rational_quadratic.py:125-127 / :197-199."""
    u_x = h[..., :n_bins]
    u_y = h[..., n_bins:2 * n_bins]
    u_d = F.pad(h[..., 2 * n_bins:], pad=(1, 1), mode='constant', value=RQ_EDGE_U)
    return u_x, u_y, u_d


def rq_forward_1d(v: Tensor, h: Tensor, n_bins: int, boundary: float):
    """rational_quadratic.py:65-128.  v:(N,), h:(N,3K-1).  Returns out, log_det, k (bin index)."""
    u_x, u_y, u_d = _rq_split(h, n_bins)
    bin_x, w = rq_bins(u_x, -boundary, boundary)
    bin_y, hg = rq_bins(u_x + u_y / 1000, -boundary, boundary)
    deltas = RQ_MIN_DELTA + F.softplus(RQ_EDGE_U + u_d / 1000)
    k = torch.searchsorted(bin_x, v[..., None]) - 1
    y_k = torch.gather(bin_y, -1, k)
    x_k = torch.gather(bin_x, -1, k)
    h_k = torch.gather(hg, -1, k)
    w_k = torch.gather(w, -1, k)
    d0 = torch.gather(deltas, -1, k)
    d1 = torch.gather(deltas, -1, k + 1)
    s = h_k / w_k
    v2 = v.view(-1, 1)
    t1 = d1 + d0 - 2 * s
    xi = (v2 - x_k) / w_k
    xi = torch.clip(xi, 0.0, 1.0)
    q = xi * (1 - xi)
    num = h_k * (s * xi ** 2 + d0 * q)
    den = s + t1 * q
    out = y_k + num / den
    ld = _rq_log_det(s, d0, d1, xi, q, t1)
    return out.view(-1), ld.view(-1), k.view(-1)


def rq_inverse_1d(v: Tensor, h: Tensor, n_bins: int, boundary: float):
    """rational_quadratic.py:130-200."""
    u_x, u_y, u_d = _rq_split(h, n_bins)
    bin_x, w = rq_bins(u_x, -boundary, boundary)
    bin_y, hg = rq_bins(u_x + u_y / 1000, -boundary, boundary)
    deltas = RQ_MIN_DELTA + F.softplus(RQ_EDGE_U + u_d / 1000)
    k = torch.searchsorted(bin_y, v[..., None]) - 1
    y_k = torch.gather(bin_y, -1, k)
    x_k = torch.gather(bin_x, -1, k)
    h_k = torch.gather(hg, -1, k)
    w_k = torch.gather(w, -1, k)
    d0 = torch.gather(deltas, -1, k)
    d1 = torch.gather(deltas, -1, k + 1)
    s = h_k / w_k
    v2 = v.view(-1, 1)
    t1 = d1 + d0 - 2 * s
    t0 = v2 - y_k
    t2 = h_k * d0
    a = (h_k * s - t2) + t0 * t1
    b = t2 - t0 * t1
    c = -s * t0
    sq = torch.clip(torch.sqrt(b ** 2 - 4 * a * c), min=0.0)
    xi = 2 * c / (-b - sq)
    xi = torch.clip(xi, 0.0, 1.0)
    q = xi * (1 - xi)
    out = xi * w_k + x_k
    ld = -_rq_log_det(s, d0, d1, xi, q, t1)
    return out.view(-1), ld.view(-1), k.view(-1)


def rq_forward(x: Tensor, h: Tensor, n_bins: int = 8, boundary: float = 50.0, n_event_dims: int = 1,
               return_bins: bool = False):
    """MonotonicSpline.forward (spline/base.py:53-60): strict in-bounds mask, identity tails."""
    z = torch.clone(x)
    ld = torch.zeros_like(z)
    kk = torch.full(x.shape, -1, dtype=torch.int64, device=x.device)
    mask = (x > -boundary) & (x < boundary)
    if torch.any(mask):
        z[mask], ld[mask], kk[mask] = rq_forward_1d(x[mask], h[mask], n_bins, boundary)
    ld = _sum_event(ld, n_event_dims)
    return (z, ld, kk) if return_bins else (z, ld)


def rq_inverse(z: Tensor, h: Tensor, n_bins: int = 8, boundary: float = 50.0, n_event_dims: int = 1,
               return_bins: bool = False):
    """MonotonicSpline.inverse (spline/base.py:65-72)."""
    x = torch.clone(z)
    ld = torch.zeros_like(x)
    kk = torch.full(z.shape, -1, dtype=torch.int64, device=z.device)
    mask = (z > -boundary) & (z < boundary)
    if torch.any(mask):
        x[mask], ld[mask], kk[mask] = rq_inverse_1d(z[mask], h[mask], n_bins, boundary)
    ld = _sum_event(ld, n_event_dims)
    return (x, ld, kk) if return_bins else (x, ld)


# ---- linear rational spline (transformers/spline/linear_rational.py:9-182) ---------------------------------
LRS_MIN_BIN = 1e-2                                        # linear_rational.py:19-20
LRS_MIN_D = 1e-5                                          # linear_rational.py:21
LRS_CONST = math.log(math.exp(1 - LRS_MIN_D) - 1)         # linear_rational.py:22
LRS_EPS = 5e-10                                           # linear_rational.py:23


def lrs_bins(u: Tensor, lo: float, hi: float, n_bins: int) -> Tensor:
    """linear_rational.py:67-75 (compute_bins)."""
    sizes = torch.softmax(u, dim=-1)
    sizes = LRS_MIN_BIN + (1 - LRS_MIN_BIN * n_bins) * sizes
    bins = torch.cumsum(sizes, dim=-1)
    bins = F.pad(bins, pad=(1, 0), mode='constant', value=0.0)
    bins = (hi - lo) * bins + lo
    bins[..., 0] = lo
    bins[..., -1] = hi
    return bins


def _lrs_knots(h: Tensor, n_bins: int, boundary: float):
    """linear_rational.py:82-90 (compute_knots) and the parameter split of :96-100."""
    K = n_bins
    u_x, u_y, u_l, u_d, u_w0 = h[:, :K], h[:, K:2 * K], h[:, 2 * K:3 * K], h[:, 3 * K:4 * K - 1], h[:, 4 * K - 1]
    knots_x = lrs_bins(u_x, -boundary, boundary, K)
    knots_y = lrs_bins(u_x + u_y / 100, -boundary, boundary, K)
    knots_lambda = torch.sigmoid(u_l)
    knots_d = F.pad(F.softplus(LRS_CONST + u_d / 100) + LRS_MIN_D, pad=(1, 1), mode='constant', value=1.0)
    return knots_x, knots_y, knots_d, knots_lambda, u_w0


def _lrs_parameters(idx, knots_x, knots_y, knots_d, knots_lambda, u_w0):
    """linear_rational.py:30-65 (compute_parameters)."""
    w0 = F.softplus(u_w0)
    w = w0[:, None] * torch.sqrt(knots_d[:, 0][:, None] / knots_d)
    w_k, w_kp1 = w.gather(1, idx), w.gather(1, idx + 1)
    lambda_k = knots_lambda.gather(1, idx)
    x_k, x_kp1 = knots_x.gather(1, idx), knots_x.gather(1, idx + 1)
    y_k, y_kp1 = knots_y.gather(1, idx), knots_y.gather(1, idx + 1)
    d_k, d_kp1 = knots_d.gather(1, idx), knots_d.gather(1, idx + 1)
    y_m = torch.divide((1 - lambda_k) * w_k * y_k + lambda_k * w_kp1 * y_kp1, (1 - lambda_k) * w_k + lambda_k * w_kp1)
    w_m = torch.multiply(lambda_k * w_k * d_k + (1 - lambda_k) * w_kp1 * d_kp1, torch.divide(x_kp1 - x_k, y_kp1 - y_k))
    return lambda_k, w_k, w_m, w_kp1, x_k, x_kp1, y_k, y_m, y_kp1


def lrs_forward_1d(x: Tensor, h: Tensor, n_bins: int, boundary: float):
    """linear_rational.py:92-135."""
    knots_x, knots_y, knots_d, knots_lambda, u_w0 = _lrs_knots(h, n_bins, boundary)
    idx = torch.searchsorted(knots_x, x[:, None]) - 1
    lam, w_k, w_m, w_kp1, x_k, x_kp1, y_k, y_m, y_kp1 = _lrs_parameters(idx, knots_x, knots_y, knots_d, knots_lambda, u_w0)
    phi = (x[:, None] - x_k) / (x_kp1 - x_k)
    mask = phi > lam
    out_lt = torch.divide(w_k * y_k * (lam - phi) + w_m * y_m * phi, w_k * (lam - phi) + w_m * phi)
    ld_lt = (torch.log(lam * w_k * w_m * (y_m - y_k)) - torch.log((w_k * (lam - phi) + w_m * phi) ** 2 + LRS_EPS)
             - torch.log(x_kp1 - x_k))
    out_gt = torch.divide(w_m * y_m * (1 - phi) + w_kp1 * y_kp1 * (phi - lam), w_m * (1 - phi) + w_kp1 * (phi - lam))
    ld_gt = (torch.log((1 - lam) * w_m * w_kp1 * (y_kp1 - y_m))
             - torch.log((w_m * (1 - phi) + w_kp1 * (phi - lam)) ** 2 + LRS_EPS) - torch.log(x_kp1 - x_k))
    return torch.where(mask, out_gt, out_lt).flatten(), torch.where(mask, ld_gt, ld_lt).flatten()


def lrs_inverse_1d(z: Tensor, h: Tensor, n_bins: int, boundary: float):
    """linear_rational.py:137-182."""
    knots_x, knots_y, knots_d, knots_lambda, u_w0 = _lrs_knots(h, n_bins, boundary)
    idx = torch.searchsorted(knots_y, z[:, None]) - 1
    lam, w_k, w_m, w_kp1, x_k, x_kp1, y_k, y_m, y_kp1 = _lrs_parameters(idx, knots_x, knots_y, knots_d, knots_lambda, u_w0)
    z = z[:, None]
    mask = z > y_m
    out_lt = torch.divide(lam * w_k * (y_k - z), w_k * (y_k - z) + w_m * (z - y_m)) * (x_kp1 - x_k) + x_k
    ld_lt = (torch.log(lam * w_k * w_m * (y_m - y_k)) - torch.log((w_k * (y_k - z) + w_m * (z - y_m)) ** 2 + LRS_EPS)
             + torch.log(x_kp1 - x_k))
    out_gt = torch.divide(lam * w_kp1 * (y_kp1 - z) + w_m * (z - y_m), w_kp1 * (y_kp1 - z) + w_m * (z - y_m)) * (x_kp1 - x_k) + x_k
    ld_gt = (torch.log((1 - lam) * w_m * w_kp1 * (y_kp1 - y_m))
             - torch.log((w_kp1 * (y_kp1 - z) + w_m * (z - y_m)) ** 2 + LRS_EPS) + torch.log(x_kp1 - x_k))
    return torch.where(mask, out_gt, out_lt).flatten(), torch.where(mask, ld_gt, ld_lt).flatten()


def _lrs_apply(fn, v: Tensor, h: Tensor, n_bins: int, boundary: float, n_event_dims: int):
    """MonotonicSpline.forward / inverse (spline/base.py:53-72): strict in-bounds mask, identity tails."""
    out = torch.clone(v)
    ld = torch.zeros_like(out)
    mask = (v > -boundary) & (v < boundary)
    if torch.any(mask):
        o, l = fn(v[mask], h[mask], n_bins, boundary)
        out = out.masked_scatter(mask, o)
        ld = ld.masked_scatter(mask, l)
    return out, _sum_event(ld, n_event_dims)


def lrs_forward(x: Tensor, h: Tensor, n_bins: int = 8, boundary: float = 50.0, n_event_dims: int = 1):
    return _lrs_apply(lrs_forward_1d, x, h, n_bins, boundary, n_event_dims)


def lrs_inverse(z: Tensor, h: Tensor, n_bins: int = 8, boundary: float = 50.0, n_event_dims: int = 1):
    return _lrs_apply(lrs_inverse_1d, z, h, n_bins, boundary, n_event_dims)


# ---- Scale (transformers/linear/affine.py:160-200): z = alpha * x, alpha as for Affine with its own constant ----------
def scale_forward(x: Tensor, h: Tensor, n_event_dims: int = 1) -> Tuple[Tensor, Tensor]:
    alpha = torch.exp(math.log(1 - AFFINE_MIN_SCALE) + h[..., 0] / 2.0) + AFFINE_MIN_SCALE
    return alpha * x, _sum_event(torch.log(alpha), n_event_dims)


def scale_inverse(z: Tensor, h: Tensor, n_event_dims: int = 1) -> Tuple[Tensor, Tensor]:
    alpha = torch.exp(math.log(1 - AFFINE_MIN_SCALE) + h[..., 0] / 2.0) + AFFINE_MIN_SCALE
    return z / alpha, -_sum_event(torch.log(alpha), n_event_dims)


TRANSFORMERS = {
    # kind -> (forward, inverse, params per element)
    'affine': (affine_forward, affine_inverse),
    'inverse_affine': (affine_inverse, affine_forward),   # affine.py:62-70
    'shift': (shift_forward, shift_inverse),
    'rq': (rq_forward, rq_inverse),
    'lrs': (lrs_forward, lrs_inverse),
    'scale': (scale_forward, scale_inverse),
}


def transformer_apply(kind: str, direction: str, x: Tensor, h: Tensor, **kw):
    fwd, inv = TRANSFORMERS[kind]
    fn = fwd if direction == 'forward' else inv
    if kind in ('rq', 'lrs'):
        return fn(x, h, **kw)
    return fn(x, h)


def params_per_element(kind: str, n_bins: int = 8) -> int:
    return {'affine': 2, 'inverse_affine': 2, 'shift': 1, 'rq': 3 * n_bins - 1, 'lrs': 4 * n_bins, 'scale': 1}[kind]


# ----------------------------------------------------------------------------------------------
# Conditioners   (conditioning/transforms.py)
# ----------------------------------------------------------------------------------------------
def feedforward(x_a: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """FeedForward with Tanh between Linear layers (transforms.py:293-307).  Layer indices
    0,2,4,... of ``sequential`` are the Linear modules."""
    idx = sorted({int(k[len(prefix):].split('.')[1]) for k in sd
                  if k.startswith(prefix + 'sequential.') and k.endswith('.weight')})
    a = x_a
    for n, i in enumerate(idx):
        a = F.linear(a, sd[f'{prefix}sequential.{i}.weight'], sd[f'{prefix}sequential.{i}.bias'])
        if n + 1 < len(idx):
            a = torch.tanh(a)
    return a


def made_degrees(n_in: int, n_hidden: int, n_out: int, n_layers: int = 2) -> List[Tensor]:
    """transforms.py:222-226."""
    return [torch.arange(n_in) + 1,
            *[(torch.arange(n_hidden) % (n_in - 1)) + 1 for _ in range(n_layers - 1)],
            torch.arange(n_out) + 1]


def made_masks(degrees: Sequence[Tensor]) -> List[Tensor]:
    """transforms.py:246-257: hidden layers use >=, the output layer uses >."""
    n_layers = len(degrees) - 1
    masks = []
    for i in range(1, n_layers + 1):
        cur, prev = degrees[i][:, None], degrees[i - 1][None, :]
        masks.append(((cur > prev) if i == n_layers else (cur >= prev)).float())
    return masks


def made(x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """MADE forward with MaskedLinear = F.linear(x, W*mask, b) (transforms.py:197-198,259-264)."""
    idx = sorted({int(k[len(prefix):].split('.')[1]) for k in sd
                  if k.startswith(prefix + 'sequential.') and k.endswith('.weight')})
    a = x
    for n, i in enumerate(idx):
        p = f'{prefix}sequential.{i}.'
        a = F.linear(a, sd[p + 'weight'] * sd[p + 'mask'], sd[p + 'bias'])
        if n + 1 < len(idx):
            a = torch.tanh(a)
    return a


# ----------------------------------------------------------------------------------------------
# Layers
# ----------------------------------------------------------------------------------------------
class LayerSpec:
    """One entry of ``bijection.layers`` (architectures.py:44-54)."""

    def __init__(self, kind: str, index: int, transformer: Optional[str] = None):
        self.kind = kind            # 'elementwise_affine' | 'actnorm' | 'reverse' | 'coupling' | 'ma' | 'inverse_ma'
        self.index = index
        self.transformer = transformer

    @property
    def prefix(self):
        return f'bijection.layers.{self.index}.'


PRESETS = {
    # preset -> (layer kind, transformer)           architectures.py:57-163, layers.py:102-395
    'NICE': ('coupling', 'shift'),
    'RealNVP': ('coupling', 'affine'),
    'InverseRealNVP': ('coupling', 'inverse_affine'),
    'MAF': ('ma', 'affine'),
    'IAF': ('inverse_ma', 'inverse_affine'),
    'CouplingRQNSF': ('coupling', 'rq'),
    'MaskedAutoregressiveRQNSF': ('ma', 'rq'),
    'InverseAutoregressiveRQNSF': ('inverse_ma', 'rq'),
    'CouplingLRS': ('coupling', 'lrs'),                   # architectures.py:166-223
    'MaskedAutoregressiveLRS': ('ma', 'lrs'),
    'InverseAutoregressiveLRS': ('inverse_ma', 'lrs'),
}


def preset_layers(preset: str, n_layers: int = 2) -> List[LayerSpec]:
    """[EA] + n_layers x [Rev, Base, ActNorm] + [EA, ActNorm]   (architectures.py:44-54)."""
    kind, tr = PRESETS[preset]
    specs = [LayerSpec('elementwise_affine', 0)]
    i = 1
    for _ in range(n_layers):
        specs.append(LayerSpec('reverse', i)); i += 1
        specs.append(LayerSpec(kind, i, tr)); i += 1
        specs.append(LayerSpec('actnorm', i)); i += 1
    specs.append(LayerSpec('elementwise_affine', i)); i += 1
    specs.append(LayerSpec('actnorm', i))
    return specs


class OracleFlow:
    """Flow(preset(event_shape)) evaluated from a reference ``state_dict`` (flows.py:606-713)."""

    def __init__(self, preset: str, event_shape, state_dict: Dict[str, Tensor], n_layers: int = 2,
                 n_bins: int = 8, boundary: float = 50.0, device: str = 'cpu'):
        """``device='cuda:0'`` runs the very same ATen op sequence eagerly on the GPU: the reference's own "CUDA support"
        is nn.Module.cuda() (docs/source/guides/cuda.rst:4-18), so this is the existing-Blackwell-path baseline of
        bench.py's ``gpu_eager_baseline``; parity checks always use the CPU default."""
        if isinstance(event_shape, int):
            event_shape = (event_shape,)
        self.preset = preset
        self.event_shape = tuple(event_shape)
        self.n_dim = int(math.prod(self.event_shape))
        self.sd = {k: v.detach().clone().to(device) for k, v in state_dict.items()}
        self.layers = preset_layers(preset, n_layers)
        self.n_bins = n_bins
        self.boundary = boundary

    # -- helpers -------------------------------------------------------------------------------
    def _tkw(self, spec):
        return dict(n_bins=self.n_bins, boundary=self.boundary) if spec.transformer in ('rq', 'lrs') else {}

    def _flat(self, x: Tensor) -> Tensor:
        return x.reshape(*x.shape[: x.dim() - len(self.event_shape)], self.n_dim)

    # -- per-layer maps (operating on event-flattened tensors (*batch, D)) ------------------------
    def _elementwise(self, spec, x, direction):
        """ElementwiseAffine (layers.py:19-26) / ActNorm = InverseAffine (layers.py:39-69); the
        global parameter ``value`` is broadcast over the batch (layers_base.py:300-303)."""
        value = self.sd[spec.prefix + 'value'].reshape(self.n_dim, 2)
        h = value.expand(*x.shape[:-1], self.n_dim, 2)
        kind = 'affine' if spec.kind == 'elementwise_affine' else 'inverse_affine'
        return transformer_apply(kind, direction, x, h)

    def _coupling(self, spec, x, direction):
        """CouplingBijection.forward/inverse with HalfSplit (layers_base.py:119-163,
        coupling_masks.py:78-81): source = first D//2 flat dims."""
        ds = self.n_dim // 2
        dt = self.n_dim - ds
        p = params_per_element(spec.transformer, self.n_bins)
        out = x.clone()
        h = feedforward(x[..., :ds], self.sd, spec.prefix + 'conditioner_transform.')
        h = h.view(*x.shape[:-1], dt, p)
        yb, ld = transformer_apply(spec.transformer, direction, x[..., ds:], h, **self._tkw(spec))
        out[..., ds:] = yb
        return out, ld

    def _ma_one_pass(self, spec, x, direction):
        """MaskedAutoregressiveBijection.apply_conditioner_transformer (layers_base.py:202-208)."""
        p = params_per_element(spec.transformer, self.n_bins)
        h = made(x, self.sd, spec.prefix + 'conditioner_transform.').view(*x.shape[:-1], self.n_dim, p)
        return transformer_apply(spec.transformer, direction, x, h, **self._tkw(spec))

    def _ma_sequential(self, spec, z):
        """MaskedAutoregressiveBijection.inverse (layers_base.py:213-223): D full passes; returns
        the log-det of the LAST pass."""
        x = torch.clone(z)
        ld = torch.zeros(z.shape[:-1], device=z.device)
        for i in range(self.n_dim):
            tmp, ld = self._ma_one_pass(spec, torch.clone(x), 'inverse')
            x[..., i] = tmp[..., i]
        return x, ld

    def layer_apply(self, spec: LayerSpec, x: Tensor, direction: str) -> Tuple[Tensor, Tensor]:
        if spec.kind in ('elementwise_affine', 'actnorm'):
            return self._elementwise(spec, x, direction)
        if spec.kind == 'reverse':  # matrix/permutation.py:19-37 (a flip is its own inverse), log-det 0
            return torch.flip(x, dims=(-1,)), torch.zeros(x.shape[:-1], device=x.device)
        if spec.kind == 'coupling':
            return self._coupling(spec, x, direction)
        if spec.kind == 'ma':
            return self._ma_one_pass(spec, x, 'forward') if direction == 'forward' else self._ma_sequential(spec, x)
        if spec.kind == 'inverse_ma':  # layers_base.py:226-234 swaps the two directions
            return self._ma_sequential(spec, x) if direction == 'forward' else self._ma_one_pass(spec, x, 'forward')
        raise ValueError(spec.kind)

    # -- composition (bijections/base.py:203-232) -------------------------------------------------
    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        xf = self._flat(x)
        log_det = torch.zeros(xf.shape[:-1], device=xf.device)
        for spec in self.layers:
            xf, ld = self.layer_apply(spec, xf, 'forward')
            log_det += ld
        return xf.view(x.shape), log_det

    def inverse(self, z: Tensor) -> Tuple[Tensor, Tensor]:
        zf = self._flat(z)
        log_det = torch.zeros(zf.shape[:-1], device=zf.device)
        for spec in self.layers[::-1]:
            zf, ld = self.layer_apply(spec, zf, 'inverse')
            log_det += ld
        return zf.view(z.shape), log_det

    # -- distribution API (flows.py:628-713, base_distributions/gaussian.py:46-54) ----------------
    def base_log_prob(self, z: Tensor) -> Tensor:
        zf = self._flat(z)
        loc, log_scale = self.sd['base.loc'], self.sd['base.log_scale']
        scale = torch.exp(log_scale)
        e = -(0.5 * ((zf - loc) / scale) ** 2 + 0.5 * math.log(2 * math.pi) + log_scale)
        return torch.sum(e, dim=-1)

    def log_prob(self, x: Tensor) -> Tensor:
        z, log_det = self.forward(x)
        return self.base_log_prob(z) + log_det

    def sample_from_noise(self, noise: Tensor, return_log_prob: bool = False):
        """Flow.sample with the base draw factored out: ``noise`` plays torch.randn of gaussian.py:42.
        NB: returns base_log_prob(z) + log_det_inverse (flows.py:710-712)."""
        z = self.sd['base.loc'] + self._flat(noise) * torch.exp(self.sd['base.log_scale'])
        z = z.view(noise.shape)
        x, log_det = self.inverse(z)
        if return_log_prob:
            return x, self.base_log_prob(z) + log_det
        return x

    # -- training-mode pieces ----------------------------------------------------------------------
    def actnorm_initialise(self, x: Tensor) -> None:
        """Data-dependent ActNorm init performed by the first training-mode forward
        (layers.py:58-68): every ActNorm sees the activations at its own depth."""
        xf = self._flat(x).reshape(-1, self.n_dim)
        for spec in self.layers:
            if spec.kind == 'actnorm':
                shift = torch.mean(xf, dim=0)[..., None]
                scale = torch.ones_like(shift) if xf.shape[0] == 1 else torch.std(xf, dim=0)[..., None]
                self.sd[spec.prefix + 'value'] = torch.cat([affine_unconstrain_scale(scale), shift], dim=-1).view(
                    *self.event_shape, 2)
            xf, _ = self.layer_apply(spec, xf, 'forward')

    def regularization(self) -> Tensor:
        """0.01 * sum theta^2 over the trainable parameters of coupling / MA layers
        (layers_base.py:38-48,75,182; bijections/base.py:234-243)."""
        total = torch.tensor(0.0)
        for spec in self.layers:
            if spec.kind in ('coupling', 'ma', 'inverse_ma'):
                pre = spec.prefix + 'conditioner_transform.'
                sq = sum(torch.sum(torch.square(v)) for k, v in self.sd.items()
                         if k.startswith(pre) and (k.endswith('weight') or k.endswith('bias')))
                total = total + 0.01 * sq
        return total

    def batch_loss(self, x: Tensor, w: Optional[Tensor] = None) -> Tensor:
        """flows.py:199-224: -mean(w * log_prob) / event_size + regularization."""
        lp = self.log_prob(x)
        if w is None:
            w = torch.ones_like(lp)
        return -torch.mean(lp * w) / self.n_dim + self.regularization()
