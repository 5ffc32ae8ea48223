"""Coupling partitions (API of torchflows/.../conditioning/coupling_masks.py:8-98).

``HalfSplit``: the first ``D // 2`` flattened event dimensions are the source (constant) part, the rest is
transformed.  The fused kernels hard-wire exactly this partition (csrc/b2f_flow.cu); other partitions go
through the composite path."""
from typing import List, Tuple

import torch

from torchflows_b200.utils import event_size


class PartialCoupling:
    def __init__(self, event_shape, source_mask: torch.Tensor, target_mask: torch.Tensor):
        self.event_shape = event_shape
        self.source_mask = source_mask
        self.target_mask = target_mask
        self.event_size = event_size(event_shape)

    @property
    def ignored_event_size(self):
        return torch.sum(1 - (self.source_mask + self.target_mask))

    @property
    def source_event_size(self) -> int:
        return int(torch.sum(self.source_mask))

    @property
    def constant_shape(self) -> Tuple[int, ...]:
        return (self.source_event_size,)

    @property
    def target_event_size(self) -> int:
        return int(torch.sum(self.target_mask))

    @property
    def target_shape(self) -> Tuple[int, ...]:
        return (self.target_event_size,)


class Coupling(PartialCoupling):
    def __init__(self, event_shape, mask: torch.Tensor):
        super().__init__(event_shape, source_mask=mask, target_mask=~mask)

    @property
    def ignored_event_size(self):
        return 0


class GraphicalCoupling(PartialCoupling):
    def __init__(self, event_shape, edge_list: List[Tuple[int, int]]):
        if len(event_shape) != 1:
            raise ValueError('GraphicalCoupling is currently only implemented for vector data')
        n = event_size(event_shape)
        sources = torch.tensor(sorted({e[0] for e in edge_list}))
        targets = torch.tensor(sorted({e[1] for e in edge_list}))
        super().__init__(event_shape, torch.isin(torch.arange(n), sources), torch.isin(torch.arange(n), targets))


class HalfSplit(Coupling):
    def __init__(self, event_shape):
        n = event_size(event_shape)
        super().__init__(event_shape, mask=(torch.arange(n) < n // 2).view(*event_shape))


def make_coupling(event_shape, edge_list: List[Tuple[int, int]] = None, coupling_type: str = 'half_split', **kwargs):
    if edge_list is not None:
        return GraphicalCoupling(event_shape, edge_list)
    if coupling_type == 'half_split':
        return HalfSplit(event_shape)
    raise ValueError(coupling_type)
