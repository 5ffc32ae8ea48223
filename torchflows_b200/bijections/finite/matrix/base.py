"""Invertible-matrix bijection base (API of torchflows/bijections/finite/matrix/base.py:9-73)."""
from typing import Tuple, Union

import torch

from torchflows_b200.bijections.base import Bijection
from torchflows_b200.utils import get_batch_shape


class InvertibleMatrix(Bijection):
    def __init__(self, event_shape: Union[torch.Size, Tuple[int, ...]], l2_regularization: bool = False, **kwargs):
        super().__init__(event_shape, **kwargs)
        self.l2_regularization = l2_regularization
        self.register_buffer('device_buffer', torch.zeros(1))

    def project_flat(self, x_flat: torch.Tensor, context_flat: torch.Tensor = None) -> torch.Tensor:
        raise NotImplementedError

    def solve_flat(self, b_flat: torch.Tensor, context: torch.Tensor = None) -> torch.Tensor:
        raise NotImplementedError

    def log_det_project(self) -> torch.Tensor:
        raise NotImplementedError

    def forward(self, x: torch.Tensor, context: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
        batch_shape = get_batch_shape(x, self.event_shape)
        z = self.project_flat(x.reshape(*batch_shape, -1)).reshape(x.shape)
        return z, self.log_det_project().to(x.device).expand(*batch_shape, 1).squeeze(-1).clone()

    def inverse(self, z: torch.Tensor, context: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
        batch_shape = get_batch_shape(z, self.event_shape)
        x = self.solve_flat(z.reshape(*batch_shape, -1)).reshape(z.shape)
        return x, -self.log_det_project().to(z.device).expand(*batch_shape, 1).squeeze(-1).clone()

    def regularization(self, *aux):
        if self.l2_regularization:
            return sum([torch.sum(torch.square(p)) for p in self.parameters() if p.requires_grad])
        return torch.tensor(0.0)
