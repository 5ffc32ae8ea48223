"""Diagonal Gaussian base distribution (API of torchflows/base_distributions/gaussian.py:8-54).

Inside ``Flow.log_prob`` / ``Flow.sample`` of a fused flow the density is evaluated by the flow kernel
itself (csrc/b2f_flow.cu, gauss_logp); the methods here are the stand-alone API and use plain torch ops on
whatever device the buffers live on."""
import math

import torch
import torch.nn as nn


class DiagonalGaussian(torch.distributions.Distribution, nn.Module):
    def __init__(self, loc: torch.Tensor, scale: torch.Tensor, trainable_loc: bool = False,
                 trainable_scale: bool = False):
        super().__init__(event_shape=loc.shape, validate_args=False)
        self.log_2_pi = math.log(2 * math.pi)
        if trainable_loc:
            self.register_parameter('loc', nn.Parameter(loc))
        else:
            self.register_buffer('loc', loc)
        if trainable_scale:
            self.register_parameter('log_scale', nn.Parameter(torch.log(scale)))
        else:
            self.register_buffer('log_scale', torch.log(scale))

    @property
    def scale(self) -> torch.Tensor:
        return torch.exp(self.log_scale)

    def sample(self, sample_shape: torch.Size = torch.Size()) -> torch.Tensor:
        """Noise is drawn directly on the distribution's device (the reference draws on the CPU and copies,
        gaussian.py:42; same distribution, different random stream)."""
        noise = torch.randn(*sample_shape, *self.event_shape, device=self.loc.device, dtype=self.loc.dtype)
        return self.loc + noise * self.scale

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        n_event = len(self.event_shape)
        if value.dim() <= n_event:
            raise ValueError('Incorrect input shape')
        e = -(0.5 * ((value - self.loc) / self.scale) ** 2 + 0.5 * self.log_2_pi + self.log_scale)
        return e.sum(dim=tuple(range(value.dim() - n_event, value.dim())))


class StandardGaussian(DiagonalGaussian):
    def __init__(self, event_shape):
        super().__init__(torch.zeros(size=event_shape), torch.ones(size=event_shape))
