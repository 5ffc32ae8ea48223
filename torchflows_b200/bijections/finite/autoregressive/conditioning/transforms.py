"""Conditioner networks (API of torchflows/.../conditioning/transforms.py:11-118,140-171,184-307).

``FeedForward`` (Linear-Tanh-Linear by default) and ``MADE`` (masked Linear-Tanh-masked Linear) keep the
reference's module tree so that ``state_dict`` keys are identical (``sequential.{0,2}.weight|bias|mask``,
``global_theta_flat``).  On the fused path these modules are only parameter containers: the flow kernels read
the weights directly and never materialise the output ``h``.  Calling a conditioner on its own
(``forward(x) -> h``) is the stand-alone API and is two library GEMMs (``F.linear``) on the module's device."""
import math
from typing import Optional, Type

import torch
import torch.nn as nn

from torchflows_b200.bijections.finite.autoregressive.conditioning.context import Concatenation, ContextCombiner
from torchflows_b200.utils import event_size, get_batch_shape


class ConditionerTransform(nn.Module):
    """theta = f(x, context) with theta.shape = (*batch, *parameter_shape)."""

    def __init__(self, input_event_shape, context_shape, parameter_shape, context_combiner: ContextCombiner = None,
                 global_parameter_mask: Optional[torch.Tensor] = None, initial_global_parameter_value: float = None,
                 output_lower_bound: float = -torch.inf, output_upper_bound: float = torch.inf, **kwargs):
        super().__init__()
        if global_parameter_mask is not None and global_parameter_mask.shape != parameter_shape:
            raise ValueError(f'Global parameter mask must have shape equal to the output parameter shape '
                             f'{parameter_shape}, but found {global_parameter_mask.shape}')
        self.output_lower_bound, self.output_upper_bound = output_lower_bound, output_upper_bound
        self.context_combiner = context_combiner or Concatenation(input_event_shape, context_shape)
        self.input_event_shape, self.context_shape = input_event_shape, context_shape
        self.n_input_event_dims = self.context_combiner.n_output_dims
        self.parameter_shape = parameter_shape
        self.global_parameter_mask = global_parameter_mask
        self.n_transformer_parameters = event_size(parameter_shape)
        self.n_global_parameters = 0 if global_parameter_mask is None else int(torch.sum(global_parameter_mask))
        self.n_predicted_parameters = self.n_transformer_parameters - self.n_global_parameters
        if initial_global_parameter_value is None:
            init = torch.randn(size=(self.n_global_parameters,))
        else:
            init = torch.full(size=(self.n_global_parameters,), fill_value=initial_global_parameter_value)
        self.global_theta_flat = nn.Parameter(init)

    @property
    def is_plain(self) -> bool:
        """True when every parameter is predicted and unbounded -- the only configuration the fused kernels
        take over (the default of every preset)."""
        return (self.n_global_parameters == 0 and self.output_lower_bound == -torch.inf
                and self.output_upper_bound == torch.inf)

    def get_batch_shape(self, x: torch.Tensor, context: torch.Tensor):
        if x is not None:
            return get_batch_shape(x, self.input_event_shape)
        if context is not None:
            return get_batch_shape(context, self.context_shape)
        raise ValueError('At least one of x or context must be provided.')

    def forward(self, x: torch.Tensor, context: torch.Tensor = None):
        batch_shape = self.get_batch_shape(x, context)
        if self.n_global_parameters == 0:
            out = self.predict_theta_flat(x, context).view(*batch_shape, *self.parameter_shape)
        else:
            ref = x if x is not None else context
            out = torch.zeros(*batch_shape, *self.parameter_shape, device=ref.device, dtype=ref.dtype)
            out[..., self.global_parameter_mask] = self.global_theta_flat
            if self.n_global_parameters < self.n_transformer_parameters:
                out[..., ~self.global_parameter_mask] = self.predict_theta_flat(x, context)
        lo, hi = self.output_lower_bound, self.output_upper_bound
        if lo > -torch.inf and hi < torch.inf:
            out = torch.sigmoid(out) * (hi - lo) + lo
        elif lo > -torch.inf:
            out = torch.exp(out) + lo
        elif hi < torch.inf:
            out = hi - torch.exp(out)
        return out

    def predict_theta_flat(self, x: torch.Tensor, context: torch.Tensor = None):
        raise NotImplementedError


class ElementwiseConditionerTransform(ConditionerTransform):
    """One parameter vector per element of the transformed tensor."""

    def __init__(self, input_event_shape, transformed_event_shape, parameter_shape_per_element, context_shape=None,
                 **kwargs):
        super().__init__(input_event_shape=input_event_shape, context_shape=context_shape,
                         parameter_shape=(*transformed_event_shape, *parameter_shape_per_element), **kwargs)


class TensorConditionerTransform(ConditionerTransform):
    """One parameter tensor for the whole transformed tensor; optionally a random subset of it is learned
    globally instead of predicted (``percentage_global_parameters``)."""

    def __init__(self, input_event_shape, parameter_shape, context_shape=None,
                 percentage_global_parameters: float = 0.0, **kwargs):
        mask = None
        if 0.0 < percentage_global_parameters <= 1.0:
            n = event_size(parameter_shape)
            chosen = torch.randperm(n)[:int(n * percentage_global_parameters)]
            mask = torch.zeros(n, dtype=torch.bool)
            mask[chosen] = True
            mask = mask.view(*parameter_shape)
        kwargs = dict(kwargs)
        kwargs['global_parameter_mask'] = mask
        super().__init__(input_event_shape=input_event_shape, parameter_shape=parameter_shape,
                         context_shape=context_shape, **kwargs)


class Constant(TensorConditionerTransform):
    def __init__(self, event_shape, parameter_shape, fill_value: float = None):
        super().__init__(input_event_shape=event_shape, parameter_shape=parameter_shape,
                         initial_global_parameter_value=fill_value,
                         global_parameter_mask=torch.ones(parameter_shape, dtype=torch.bool))


class MADE(ElementwiseConditionerTransform):
    """Masked autoencoder for distribution estimation.  Degrees: inputs 1..n_in, hidden units
    (j mod (n_in-1)) + 1, outputs 1..n_out; hidden masks use >=, the output mask uses > and is repeated for the
    P parameters of each element (transforms.py:222-257)."""

    class MaskedLinear(nn.Linear):
        def __init__(self, in_features: int, out_features: int, mask: torch.Tensor):
            super().__init__(in_features=in_features, out_features=out_features)
            self.register_buffer('mask', mask)

        def forward(self, x):
            return nn.functional.linear(x, self.weight * self.mask, self.bias)

    def __init__(self, input_event_shape, transformed_event_shape, parameter_shape_per_element, context_shape=None,
                 n_hidden: int = None, n_layers: int = 2, **kwargs):
        super().__init__(input_event_shape=input_event_shape, transformed_event_shape=transformed_event_shape,
                         parameter_shape_per_element=parameter_shape_per_element, context_shape=context_shape,
                         **kwargs)
        P = event_size(parameter_shape_per_element)
        n_in, n_out = self.n_input_event_dims, event_size(transformed_event_shape)
        if n_hidden is None:
            n_hidden = max(int(3 * math.log10(n_in)), 4)
        self.n_hidden, self.n_layers = n_hidden, n_layers
        degrees = [torch.arange(n_in) + 1]
        degrees += [(torch.arange(n_hidden) % (n_in - 1)) + 1 for _ in range(n_layers - 1)]
        degrees += [torch.arange(n_out) + 1]
        masks = self.create_masks(n_layers, degrees)
        modules = []
        for mask in masks[:-1]:
            modules += [self.MaskedLinear(mask.shape[1], mask.shape[0], mask), nn.Tanh()]
        last = masks[-1]
        modules.append(self.MaskedLinear(last.shape[1], last.shape[0] * P, torch.repeat_interleave(last, P, dim=0)))
        self.sequential = nn.Sequential(*modules)

    @staticmethod
    def create_masks(n_layers, ms):
        masks = []
        for i in range(1, n_layers + 1):
            cur, prev = ms[i][:, None], ms[i - 1][None, :]
            masks.append(((cur > prev) if i == n_layers else (cur >= prev)).to(torch.float))
        return masks

    def finalisation_steps(self) -> torch.Tensor:
        """For the incremental sequential inverse: hidden unit j has all of its inputs once dimensions
        0..s_j-1 are known, with s_j = number of inputs its mask row lets through."""
        return self.sequential[0].mask.sum(dim=1).to(torch.int32)

    def predict_theta_flat(self, x: torch.Tensor, context: torch.Tensor = None):
        theta = self.sequential(self.context_combiner(x, context))
        if self.global_parameter_mask is None:
            return theta          # already flat: (*batch, n_out * P)
        return theta[..., ~self.global_parameter_mask]


class LinearMADE(MADE):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, n_layers=1, **kwargs)


class FeedForward(TensorConditionerTransform):
    """n_layers Linear layers with a nonlinearity between them; default hidden width
    max(int(5*log10(max(n_in, n_out))), 4) (transforms.py:290-291)."""

    def __init__(self, input_event_shape, parameter_shape, context_shape=None, n_hidden: int = None, n_layers: int = 2,
                 nonlinearity: Type[nn.Module] = nn.Tanh, **kwargs):
        super().__init__(input_event_shape=input_event_shape, context_shape=context_shape,
                         parameter_shape=parameter_shape, **kwargs)
        n_in, n_out = self.n_input_event_dims, self.n_predicted_parameters
        if n_hidden is None:
            n_hidden = max(int(5 * math.log10(max(n_in, n_out))), 4)
        if n_layers < 1:
            raise ValueError
        self.n_hidden, self.n_layers, self.nonlinearity = n_hidden, n_layers, nonlinearity
        widths = [n_in] + [n_hidden] * (n_layers - 1) + [n_out]
        modules = []
        for i in range(n_layers):
            modules.append(nn.Linear(widths[i], widths[i + 1]))
            if i + 1 < n_layers:
                modules.append(nonlinearity())
        modules.append(nn.Unflatten(dim=-1, unflattened_size=(n_out,)))
        self.sequential = nn.Sequential(*modules)

    def predict_theta_flat(self, x: torch.Tensor, context: torch.Tensor = None):
        return self.sequential(self.context_combiner(x, context))


class Linear(FeedForward):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs, n_layers=1)


class ResidualFeedForward(TensorConditionerTransform):
    """Linear -> act -> (n_layers - 2) residual blocks ``x + MLP(x)`` of ``block_size`` Linear layers -> Linear
    (API and state_dict layout of transforms.py:315-362).  Couplings that use it run as a composite: this network on the
    library's GEMMs, the transformer as its stand-alone kernel (the whole-flow kernels implement the two-layer FeedForward)."""

    class ResidualBlock(nn.Module):
        def __init__(self, event_size: int, hidden_size: int, block_size: int, nonlinearity: Type[nn.Module]):
            super().__init__()
            if block_size < 2:
                raise ValueError(f'block_size must be at least 2 but found {block_size}. '
                                 f'For block_size = 1, use the FeedForward class instead.')
            widths = [event_size] + [hidden_size] * (block_size - 1) + [event_size]
            modules = []
            for i in range(block_size):
                modules.append(nn.Linear(widths[i], widths[i + 1]))
                if i + 1 < block_size:
                    modules.append(nonlinearity())
            self.sequential = nn.Sequential(*modules)

        def forward(self, x):
            return x + self.sequential(x)

    def __init__(self, input_event_shape, parameter_shape, context_shape=None, n_hidden: int = None, n_layers: int = 3,
                 block_size: int = 2, nonlinearity: Type[nn.Module] = nn.ReLU, **kwargs):
        super().__init__(input_event_shape=input_event_shape, context_shape=context_shape,
                         parameter_shape=parameter_shape, **kwargs)
        n_in, n_out = self.n_input_event_dims, self.n_predicted_parameters
        if n_hidden is None:
            n_hidden = max(int(5 * math.log10(max(n_in, n_out))), 4)
        if n_layers <= 2:
            raise ValueError(f'Number of layers in ResidualFeedForward must be at least 3, but found {n_layers}')
        self.n_hidden, self.n_layers, self.nonlinearity = n_hidden, n_layers, nonlinearity
        modules = [nn.Linear(n_in, n_hidden), nonlinearity()]
        modules += [self.ResidualBlock(n_hidden, n_hidden, block_size, nonlinearity) for _ in range(n_layers - 2)]
        modules += [nn.Linear(n_hidden, n_out), nn.Unflatten(dim=-1, unflattened_size=(n_out,))]
        self.sequential = nn.Sequential(*modules)

    def predict_theta_flat(self, x: torch.Tensor, context: torch.Tensor = None):
        return self.sequential(self.context_combiner(x, context))
