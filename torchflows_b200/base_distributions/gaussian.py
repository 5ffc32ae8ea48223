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
        gaussian.py:42; same distribution, different random stream): on a GPU by the library's Philox kernel, with trainable
        base parameters under autograd (reparameterised draw) or on the CPU by torch.randn."""
        if (self.loc.is_cuda and self.loc.dtype == torch.float32
                and not (torch.is_grad_enabled() and (self.loc.requires_grad or self.log_scale.requires_grad))):
            # the library's counter-based stream (csrc/b2f_philox.cuh, keyed by torch's seed): the same draws Flow.sample's
            # fused kernels make in registers
            from torchflows_b200 import _native as N
            from torchflows_b200 import _program as prog
            n, D = 1, 1
            for d in sample_shape:
                n *= int(d)
            for d in self.event_shape:
                D *= int(d)
            seed, offset = prog.next_noise_stream(self.loc.device, n * D)
            z = N.philox_normal(n, D, self.loc.device, seed, offset, self.loc.detach().reshape(-1).contiguous(),
                                self.log_scale.detach().reshape(-1).contiguous())
            return z.reshape(*sample_shape, *self.event_shape)
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
