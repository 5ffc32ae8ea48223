"""Concrete layers = (layer base, transformer) pairings (API of torchflows/.../autoregressive/layers.py)."""
from typing import Tuple, Union

import torch

from torchflows_b200 import _native as N
from torchflows_b200.bijections.finite.autoregressive.layers_base import (CouplingBijection, ElementwiseBijection,
                                                                           InverseMaskedAutoregressiveBijection,
                                                                           MaskedAutoregressiveBijection)
from torchflows_b200.bijections.finite.autoregressive.transformers.linear.affine import Affine, InverseAffine, Scale, Shift
from torchflows_b200.bijections.finite.autoregressive.transformers.spline.linear_rational import LinearRational
from torchflows_b200.bijections.finite.autoregressive.transformers.spline.rational_quadratic import RationalQuadratic


class ElementwiseAffine(ElementwiseBijection):
    def __init__(self, event_shape, **kwargs):
        super().__init__(event_shape, Affine, **kwargs)


class ElementwiseInverseAffine(ElementwiseBijection):
    def __init__(self, event_shape, **kwargs):
        super().__init__(event_shape, InverseAffine, **kwargs)


class ElementwiseScale(ElementwiseBijection):
    def __init__(self, event_shape, **kwargs):
        super().__init__(event_shape, Scale, **kwargs)


class ElementwiseShift(ElementwiseBijection):
    def __init__(self, event_shape, **kwargs):
        super().__init__(event_shape, Shift, **kwargs)


class ElementwiseRQSpline(ElementwiseBijection):
    def __init__(self, event_shape, **kwargs):
        super().__init__(event_shape, RationalQuadratic, **kwargs)


class ActNorm(ElementwiseInverseAffine):
    """z = (x - shift) / scale with frozen parameters that are initialised from the first batch seen in
    training mode: shift = batch mean, scale = unbiased batch std (layers.py:39-69).  The statistics come from
    one column-reduction kernel (b2f_column_stats, fp64 accumulation)."""

    def __init__(self, event_shape, **kwargs):
        kwargs['context_shape'] = None
        super().__init__(event_shape, **kwargs)
        self.first_training_batch_pass: bool = True
        self.value.requires_grad_(False)

    def needs_data_init(self) -> bool:
        return self.training and self.first_training_batch_pass

    @torch.no_grad()
    def data_init(self, x: torch.Tensor, reduce_fn=None):
        """``reduce_fn(sum, sumsq, count) -> (sum, sumsq, count)`` lets data-parallel training all-reduce the
        sufficient statistics so that every rank initialises identically."""
        self.first_training_batch_pass = False
        x2 = N.require_cuda_f32(x.detach(), 'ActNorm input').reshape(-1, self.n_dim)
        s, q = N.column_stats(x2)
        n = torch.tensor(float(x2.shape[0]), device=x2.device, dtype=torch.float64)
        if reduce_fn is not None:
            s, q, n = reduce_fn(s, q, n)
        mean = s / n
        if float(n) <= 1:
            std = torch.ones_like(mean)          # unit scale if it cannot be estimated (layers.py:63-64)
        else:
            std = torch.sqrt(torch.clamp((q - n * mean * mean) / (n - 1), min=0.0))
        scale = std.to(self.value.dtype)[:, None]
        shift = mean.to(self.value.dtype)[:, None]
        new = torch.cat([self.transformer.unconstrain_scale(scale), shift], dim=-1)
        with torch.no_grad():      # in place (the reference rebinds .data, layers.py:68): the parameter keeps its storage, so
            self.value.copy_(new.view(self.value.shape))      # snapshots and captured graphs keep pointing at it

    def forward(self, x: torch.Tensor, context: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.needs_data_init():
            self.data_init(x)
        return super().forward(x, context)


class AffineCoupling(CouplingBijection):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        if tuple(event_shape) == (1,):
            raise ValueError
        super().__init__(event_shape, Affine, **kwargs)


class InverseAffineCoupling(CouplingBijection):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        if tuple(event_shape) == (1,):
            raise ValueError
        super().__init__(event_shape, InverseAffine, **kwargs)


class ShiftCoupling(CouplingBijection):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, Shift, **kwargs)


class LRSCoupling(CouplingBijection):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, LinearRational, **kwargs)


class RQSCoupling(CouplingBijection):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, RationalQuadratic, **kwargs)


class AffineForwardMaskedAutoregressive(MaskedAutoregressiveBijection):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, Affine, **kwargs)


class RQSForwardMaskedAutoregressive(MaskedAutoregressiveBijection):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, RationalQuadratic, **kwargs)


class LRSForwardMaskedAutoregressive(MaskedAutoregressiveBijection):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, LinearRational, **kwargs)


class AffineInverseMaskedAutoregressive(InverseMaskedAutoregressiveBijection):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, InverseAffine, **kwargs)


class RQSInverseMaskedAutoregressive(InverseMaskedAutoregressiveBijection):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, RationalQuadratic, **kwargs)


class LRSInverseMaskedAutoregressive(InverseMaskedAutoregressiveBijection):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, LinearRational, **kwargs)


# Couplings whose name promises a one-layer conditioner.  As in the reference (layers.py:298-335) the ``n_layers=1`` keyword is
# passed to the layer, not to the conditioner, where it is swallowed: the conditioner stays the default two-layer FeedForward.
class LinearAffineCoupling(AffineCoupling):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, **kwargs, n_layers=1)


class LinearRQSCoupling(RQSCoupling):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, **kwargs, n_layers=1)


class LinearLRSCoupling(LRSCoupling):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, **kwargs, n_layers=1)


class LinearShiftCoupling(ShiftCoupling):
    def __init__(self, event_shape: Union[Tuple[int, ...], torch.Size], **kwargs):
        super().__init__(event_shape, **kwargs, n_layers=1)
