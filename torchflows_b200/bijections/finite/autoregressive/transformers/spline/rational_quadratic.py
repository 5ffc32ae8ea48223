"""Rational-quadratic spline transformer of Durkan et al. 2019 (API of
torchflows/.../transformers/spline/rational_quadratic.py:10-200).

Per element: u_x = h[:K] (width logits), u_y = h[K:2K] (heights are parameterised as u_x + u_y/1000),
u_d = h[2K:] (interior derivative logits, /1000, edges padded with log(expm1(1-1e-5))).  The whole chain
(two softmaxes, cumulative sums, knot search, rational-quadratic evaluation or its analytic inverse, log-det)
is one kernel (csrc/b2f_transformer.cu, math in csrc/b2f_math.cuh) instead of ~60 ATen launches."""
import math
from typing import Tuple, Union

import torch

from torchflows_b200 import _native as N
from torchflows_b200.bijections.finite.autoregressive.transformers.spline.base import MonotonicSpline


class RationalQuadratic(MonotonicSpline):
    _tkind_forward = N.T_RQ_FWD
    _tkind_inverse = N.T_RQ_INV

    def __init__(self, event_shape: Union[torch.Size, Tuple[int, ...]], boundary: float = 50.0, **kwargs):
        super().__init__(event_shape, min_input=-boundary, max_input=boundary, min_output=-boundary,
                         max_output=boundary, **kwargs)
        if not 1 <= self.n_bins <= 64:
            raise NotImplementedError('RationalQuadratic kernels support 1 <= n_bins <= 64')
        self.boundary = float(boundary)
        self.min_bin_size = 1e-3
        self.min_delta = 1e-5
        self.boundary_u_delta = math.log(math.expm1(1 - self.min_delta))

    @property
    def parameter_shape_per_element(self) -> torch.Size:
        return torch.Size((3 * self.n_bins - 1,))

    @property
    def default_parameters(self) -> torch.Tensor:
        return torch.zeros(self.parameter_shape)

    def _kernel_args(self):
        return self.n_bins, self.boundary

    def bin_indices(self, x: torch.Tensor, h: torch.Tensor, inverse: bool = False) -> torch.Tensor:
        """Bin index used for every element (-1 outside the bounds) -- what the reference gets from
        ``searchsorted(bin_x | bin_y, v) - 1`` (rational_quadratic.py:82,147)."""
        E, P = self.n_dim, self.n_parameters_per_element
        x2 = N.require_cuda_f32(x, 'input').reshape(-1, E)
        h3 = N.require_cuda_f32(h, 'parameters').reshape(-1, E, P)
        _, _, k = N.transformer_apply(N.T_RQ_INV if inverse else N.T_RQ_FWD, x2, h3, E * P, self.n_bins,
                                      self.boundary, want_bins=True)
        return k.reshape(x.shape)
