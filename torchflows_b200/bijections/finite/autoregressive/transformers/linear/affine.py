"""Affine / InverseAffine / Shift transformers (API of torchflows/.../transformers/linear/affine.py:10-70,137-159).

alpha = exp(log(1-m) + u0/2) + m with m = min_scale = 1e-10; z = alpha*x + u1; log_det = sum log alpha."""
import math
from typing import Tuple, Union

import torch

from torchflows_b200 import _native as N
from torchflows_b200.bijections.finite.autoregressive.transformers.base import ScalarTransformer


class Affine(ScalarTransformer):
    _tkind_forward = N.T_AFFINE_FWD
    _tkind_inverse = N.T_AFFINE_INV

    def __init__(self, event_shape: torch.Size, min_scale: float = 1e-10):
        super().__init__(event_shape=event_shape)
        if min_scale != 1e-10:
            raise NotImplementedError('the fused kernels implement min_scale = 1e-10 (the reference default)')
        self.m = min_scale
        self.identity_unconstrained_alpha = math.log(1 - self.m)
        self.const = 2

    @property
    def parameter_shape_per_element(self):
        return (2,)

    @property
    def default_parameters(self) -> torch.Tensor:
        return torch.zeros(self.parameter_shape)

    def constrain_scale(self, unconstrained_scale: torch.Tensor) -> torch.Tensor:
        return torch.exp(self.identity_unconstrained_alpha + unconstrained_scale / self.const) + self.m

    def unconstrain_scale(self, scale: torch.Tensor) -> torch.Tensor:
        return (torch.log(scale - self.m) - self.identity_unconstrained_alpha) * self.const


class InverseAffine(Affine):
    """Affine with forward and inverse swapped (affine.py:62-70)."""
    _tkind_forward = N.T_AFFINE_INV
    _tkind_inverse = N.T_AFFINE_FWD

    def __init__(self, event_shape: Union[torch.Size, Tuple[int, ...]], **kwargs):
        super().__init__(event_shape, **kwargs)


class Shift(ScalarTransformer):
    _tkind_forward = N.T_SHIFT_ADD
    _tkind_inverse = N.T_SHIFT_SUB

    def __init__(self, event_shape: torch.Size, **kwargs):
        super().__init__(event_shape=event_shape)

    @property
    def parameter_shape_per_element(self):
        return (1,)

    @property
    def default_parameters(self) -> torch.Tensor:
        return torch.zeros(self.parameter_shape)


class Scale(ScalarTransformer):
    """z = alpha * x with alpha = exp(log(1 - m) + u / 2) + m > 0 (affine.py:160-200): the Affine kernel without a shift."""
    _tkind_forward = N.T_SCALE_FWD
    _tkind_inverse = N.T_SCALE_INV

    def __init__(self, event_shape: torch.Size, min_scale: float = 1e-10):
        super().__init__(event_shape=event_shape)
        if min_scale != 1e-10:
            raise NotImplementedError('the kernels implement the default min_scale = 1e-10')
        self.m = min_scale
        self.const = 2.0
        self.u_alpha_1 = math.log(1 - self.m)

    @property
    def parameter_shape_per_element(self):
        return (1,)

    @property
    def default_parameters(self) -> torch.Tensor:
        return torch.zeros(self.parameter_shape)

    def unconstrain_alpha(self, a):
        return self.const * (torch.log(a - self.m) - self.u_alpha_1)

    def constrain_alpha(self, u):
        return torch.exp(self.u_alpha_1 + u / self.const) + self.m
