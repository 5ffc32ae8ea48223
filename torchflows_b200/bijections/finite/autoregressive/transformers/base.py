"""Transformer base classes (API of torchflows/.../transformers/base.py:8-83).

A transformer maps ``x:(*batch, *event)`` with parameters ``h:(*batch, *parameter_shape)`` to
``(z, log_det)``.  The elementwise ones below run as one sm_100a kernel launch (csrc/b2f_transformer.cu)
behind ``ElementwiseTransformerFunction``; ``h`` is element-major / parameter-minor, exactly the layout the
conditioner's last Linear produces (layers_base.py:143)."""
from typing import Tuple, Union

import torch

from torchflows_b200 import _native as N
from torchflows_b200.bijections.base import Bijection
from torchflows_b200.utils import event_size, get_batch_shape


class TensorTransformer(Bijection):
    def __init__(self, event_shape: Union[torch.Size, Tuple[int, ...]], **kwargs):
        super().__init__(event_shape=event_shape)

    def forward(self, x: torch.Tensor, h: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    def inverse(self, x: torch.Tensor, h: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    @property
    def parameter_shape(self) -> Union[torch.Size, Tuple[int, ...]]:
        raise NotImplementedError

    @property
    def n_parameters(self) -> int:
        return event_size(self.parameter_shape)

    @property
    def default_parameters(self) -> torch.Tensor:
        """Parameters of the identity map."""
        raise NotImplementedError


class ElementwiseTransformerFunction(torch.autograd.Function):
    """out, log_det = T(x; h) for one of the b2f_transformer kinds; backward by b2f_transformer_backward."""

    @staticmethod
    def forward(ctx, x2, h3, tkind, n_bins, boundary):
        # h3: (n_rows, E, P), or (1, E, P) = one parameter set broadcast over all rows (row stride 0)
        stride = 0 if h3.shape[0] == 1 and x2.shape[0] != 1 else h3.shape[1] * h3.shape[2]
        out, ld, _ = N.transformer_apply(tkind, x2, h3, stride, n_bins, boundary)
        ctx.save_for_backward(x2, h3)
        ctx.meta = (tkind, n_bins, boundary, stride)
        return out, ld

    @staticmethod
    def backward(ctx, gout, gld):
        x2, h3 = ctx.saved_tensors
        tkind, n_bins, boundary, stride = ctx.meta
        gout = None if gout is None else gout.contiguous()
        gld = None if gld is None else gld.contiguous()
        gx, gh = N.transformer_backward(tkind, x2, h3, stride, gout, gld, n_bins, boundary)
        if gh.shape[0] != h3.shape[0]:
            gh = gh.sum(dim=0, keepdim=True)
        return gx, gh, None, None, None


class ScalarTransformer(TensorTransformer):
    """Transforms every element of the event independently with its own parameter vector."""

    # kernel kinds of forward / inverse; set by subclasses
    _tkind_forward: int = -1
    _tkind_inverse: int = -1

    def __init__(self, event_shape: Union[torch.Size, Tuple[int, ...]], **kwargs):
        super().__init__(event_shape, **kwargs)

    @property
    def parameter_shape_per_element(self):
        raise NotImplementedError

    @property
    def n_parameters_per_element(self) -> int:
        return event_size(self.parameter_shape_per_element)

    @property
    def parameter_shape(self) -> Union[torch.Size, Tuple[int, ...]]:
        return torch.Size((*self.event_shape, *self.parameter_shape_per_element))

    def _kernel_args(self):
        """(n_bins, boundary) for the kernel; only splines use them."""
        return 8, 50.0

    def _apply_kernel(self, x: torch.Tensor, h: torch.Tensor, tkind: int):
        batch_shape = get_batch_shape(x, self.event_shape)
        E, P = self.n_dim, self.n_parameters_per_element
        x2 = N.require_cuda_f32(x, 'transformer input').reshape(-1, E)
        n_batch = len(batch_shape)
        if n_batch > 0 and h.dim() == n_batch + len(self.parameter_shape) and all(s == 0 for s in h.stride()[:n_batch]) \
                and x2.shape[0] > 1:
            # parameters broadcast over the batch (ElementwiseBijection.prepare_h): hand the kernel ONE copy
            h3 = h[(0,) * n_batch].reshape(1, E, P)
            if not h3.is_cuda or h3.dtype != torch.float32:
                raise N.B2FError('transformer parameters must be CUDA float32')
            h3 = h3.contiguous()
        else:
            h3 = N.require_cuda_f32(h, 'transformer parameters').reshape(-1, E, P)
            if h3.shape[0] != x2.shape[0]:
                raise ValueError(f'parameter batch {tuple(h.shape)} does not match input batch {tuple(x.shape)}')
        n_bins, boundary = self._kernel_args()
        if torch.is_grad_enabled() and (x2.requires_grad or h3.requires_grad):
            out, ld = ElementwiseTransformerFunction.apply(x2, h3, tkind, n_bins, boundary)
        else:
            stride = 0 if h3.shape[0] == 1 and x2.shape[0] != 1 else E * P
            out, ld, _ = N.transformer_apply(tkind, x2, h3, stride, n_bins, boundary)
        return out.reshape(x.shape), ld.reshape(batch_shape)

    def forward(self, x: torch.Tensor, h: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._apply_kernel(x, h, self._tkind_forward)

    def inverse(self, z: torch.Tensor, h: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._apply_kernel(z, h, self._tkind_inverse)
