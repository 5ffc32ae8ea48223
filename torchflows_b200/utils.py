"""Shape helpers and the training data loader (API of torchflows/utils.py:37-58,86,158,189-220)."""
from typing import Optional, Tuple, Union

import torch
from torch.utils.data import DataLoader, TensorDataset

Shape = Union[torch.Size, Tuple[int, ...]]


def get_batch_shape(x: torch.Tensor, event_shape: Shape) -> torch.Size:
    return x.shape[:x.dim() - len(event_shape)]


def flatten_event(x: torch.Tensor, event_shape: Shape) -> torch.Tensor:
    """(*batch, *event) -> (*batch, prod(event)).  The event size is spelled out (a -1 is ambiguous for an empty batch)."""
    n = 1
    for s in event_shape:
        n *= int(s)
    return x.reshape(*get_batch_shape(x, event_shape), n)


def unflatten_event(x: torch.Tensor, event_shape: Shape) -> torch.Tensor:
    """(*batch, prod(event)) -> (*batch, *event)."""
    return x.reshape(*x.shape[:-1], *event_shape)


def flatten_batch(x: torch.Tensor, batch_shape: Shape) -> torch.Tensor:
    return x.reshape(-1, *x.shape[len(batch_shape):])


def unflatten_batch(x: torch.Tensor, batch_shape: Shape) -> torch.Tensor:
    return x.reshape(*batch_shape, *x.shape[1:])


def sum_except_batch(x: torch.Tensor, event_shape: Shape) -> torch.Tensor:
    return x.sum(dim=tuple(range(x.dim() - len(event_shape), x.dim())))


def event_size(event_shape: Shape) -> int:
    n = 1
    for s in event_shape:
        n *= int(s)
    return n


def create_data_loader(x: torch.Tensor, weights: Optional[torch.Tensor], context: Optional[torch.Tensor], label: str,
                       event_shape: Shape, **kwargs) -> DataLoader:
    """TensorDataset + DataLoader over (x, w[, context]) -- same contract as the reference
    (torchflows/utils.py:189-220): default weights are ones, length mismatches raise ValueError."""
    if label not in ('training', 'validation', 'testing'):
        raise AssertionError(label)
    if weights is None:
        weights = torch.ones(get_batch_shape(x, event_shape))
    if len(x) != len(weights):
        raise ValueError(f'Expected same number of {label} data and {label} weights, '
                         f'but found {len(x)} and {len(weights)}')
    tensors = [x, weights]
    if context is not None:
        if len(x) != len(context):
            raise ValueError(f'Expected same number of {label} data and {label} contexts, '
                             f'but found {len(x)} and {len(context)}')
        tensors.append(context)
    return DataLoader(TensorDataset(*tensors), **kwargs)
