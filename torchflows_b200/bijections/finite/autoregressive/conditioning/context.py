"""How a conditioner input is combined with a context tensor (API of
torchflows/.../conditioning/context.py:7-64).  Only ``context=None`` is on the fused path."""
import torch
import torch.nn as nn

from torchflows_b200.utils import event_size, flatten_event


class ContextCombiner(nn.Module):
    def __init__(self, input_shape, context_shape):
        super().__init__()
        self.input_shape = input_shape
        self.context_shape = context_shape
        self.n_input_dims = event_size(input_shape) if input_shape is not None else 0
        self.n_context_dims = event_size(context_shape) if context_shape is not None else 0

    def forward(self, x: torch.Tensor, context: torch.Tensor):
        raise NotImplementedError

    @property
    def n_output_dims(self) -> int:
        raise NotImplementedError


class Concatenation(ContextCombiner):
    """cat([flatten(x), flatten(context)], -1); either part may be absent."""

    def forward(self, x: torch.Tensor, context: torch.Tensor):
        parts = []
        if x is not None:
            if self.input_shape is None:
                raise ValueError('input given but input_shape is None')
            parts.append(flatten_event(x, self.input_shape))
        if context is not None:
            if self.context_shape is None:
                raise ValueError('context given but context_shape is None')
            parts.append(flatten_event(context, self.context_shape))
        if not parts:
            raise ValueError('At least one of x or context must be provided.')
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)

    @property
    def n_output_dims(self) -> int:
        return self.n_input_dims + self.n_context_dims
