"""Monotonic spline base (API of torchflows/.../transformers/spline/base.py:9-72): identity outside
(min_input, max_input), n_bins bins inside."""
from typing import Tuple, Union

import torch

from torchflows_b200.bijections.finite.autoregressive.transformers.base import ScalarTransformer


class MonotonicSpline(ScalarTransformer):
    def __init__(self, event_shape: Union[torch.Size, Tuple[int, ...]], min_input: float = -1.0,
                 max_input: float = 1.0, min_output: float = -1.0, max_output: float = 1.0, n_bins: int = 8):
        super().__init__(event_shape)
        self.min_input, self.max_input = min_input, max_input
        self.min_output, self.max_output = min_output, max_output
        self.n_bins = n_bins
        self.n_knots = n_bins + 1

    def forward_inputs_inside_bounds_mask(self, x):
        return (x > self.min_input) & (x < self.max_input)

    def inverse_inputs_inside_bounds_mask(self, z):
        return (z > self.min_output) & (z < self.max_output)
