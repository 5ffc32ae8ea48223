"""Permutation bijections (API of torchflows/bijections/finite/matrix/permutation.py:8-37).

``ReversePermutationMatrix`` is free inside a fused flow program: it only toggles the column addressing of the
tile that the kernel keeps in shared memory (B2F_OP_FLIP)."""
from typing import Tuple, Union

import torch

from torchflows_b200 import _native as N
from torchflows_b200 import _program as prog
from torchflows_b200.bijections.finite.matrix.base import InvertibleMatrix
from torchflows_b200.utils import event_size


class PermutationMatrix(InvertibleMatrix):
    def __init__(self, event_shape: Union[torch.Size, Tuple[int, ...]], forward_permutation: torch.Tensor, **kwargs):
        super().__init__(event_shape, **kwargs)
        assert forward_permutation.shape == tuple(event_shape)
        self.forward_permutation = forward_permutation.reshape(-1)
        self.inverse_permutation = torch.empty_like(self.forward_permutation)
        self.inverse_permutation[self.forward_permutation] = torch.arange(self.n_dim)

    def project_flat(self, x_flat: torch.Tensor, context_flat: torch.Tensor = None) -> torch.Tensor:
        return x_flat[..., self.forward_permutation.to(x_flat.device)]

    def solve_flat(self, b_flat: torch.Tensor, context: torch.Tensor = None) -> torch.Tensor:
        return b_flat[..., self.inverse_permutation.to(b_flat.device)]

    def log_det_project(self) -> torch.Tensor:
        return torch.zeros(1).to(self.device_buffer.device)


class RandomPermutationMatrix(PermutationMatrix):
    def __init__(self, event_shape: Union[torch.Size, Tuple[int, ...]], **kwargs):
        n = event_size(event_shape)
        super().__init__(event_shape, forward_permutation=torch.randperm(n).view(*event_shape), **kwargs)


class ReversePermutationMatrix(PermutationMatrix):
    def __init__(self, event_shape: Union[torch.Size, Tuple[int, ...]], **kwargs):
        n = event_size(event_shape)
        super().__init__(event_shape, forward_permutation=torch.arange(n - 1, -1, -1).view(*event_shape), **kwargs)

    def lower(self, direction: str):
        from torchflows_b200.bijections.finite.autoregressive.layers_base import _fits_fused_kernel
        if not _fits_fused_kernel(self.n_dim, 1):
            return None          # very wide events: plain index flip (InvertibleMatrix.forward / inverse)
        return [prog.LoweredOp(kind=N.OP_FLIP, owner=self)]

    def column_op(self, direction: str):
        """As one op of a per-column run (csrc/b2f_colrun.cu) when the event is too wide for the whole-flow kernels."""
        return (N.COL_FLIP, None) if self.n_dim % 4 == 0 else None

    def project_flat(self, x_flat: torch.Tensor, context_flat: torch.Tensor = None) -> torch.Tensor:
        return torch.flip(x_flat, dims=(-1,))       # same map as indexing with the reversed permutation, cheap backward

    def solve_flat(self, b_flat: torch.Tensor, context: torch.Tensor = None) -> torch.Tensor:
        return torch.flip(b_flat, dims=(-1,))

    def forward(self, x, context=None):
        return self._run_fused(x, 'forward') if self.lower('forward') is not None else super().forward(x, context)

    def inverse(self, z, context=None):
        return self._run_fused(z, 'inverse') if self.lower('inverse') is not None else super().inverse(z, context)
