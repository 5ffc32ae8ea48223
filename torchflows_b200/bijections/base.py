"""Bijection contract and sequential composition (API of torchflows/bijections/base.py:11-243).

``forward(x, context=None) -> (z, log_det)`` and ``inverse(z, context=None) -> (x, log_det)`` with
``x:(*batch, *event)``, ``log_det:(*batch)``; inputs are never mutated.  New here: ``lower(direction)`` lets a
layer describe itself as ops of a libb2f flow program, and ``BijectiveComposition`` runs a whole stack of such
layers as ONE kernel launch instead of the reference's per-layer Python loop (base.py:211-222)."""
from typing import Any, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from torchflows_b200 import _program as prog
from torchflows_b200.utils import event_size, get_batch_shape


class Bijection(nn.Module):
    def __init__(self, event_shape: Union[torch.Size, Tuple[int, ...]],
                 context_shape: Union[torch.Size, Tuple[int, ...]] = None, **kwargs):
        super().__init__()
        self.event_shape = event_shape
        self.n_dim = event_size(event_shape)
        self.context_shape = context_shape

    def forward(self, x: torch.Tensor, context: torch.Tensor = None) -> Tuple[torch.Tensor, ...]:
        raise NotImplementedError

    def inverse(self, z: torch.Tensor, context: torch.Tensor = None) -> Tuple[torch.Tensor, ...]:
        raise NotImplementedError

    # -- fused-path hook ---------------------------------------------------------------------------
    def lower(self, direction: str) -> Optional[List[prog.LoweredOp]]:
        """Ops of this layer for direction 'forward' | 'inverse', or None if it cannot join a fused program."""
        return None

    def _run_fused(self, x: torch.Tensor, direction: str):
        ops = self.lower(direction)
        batch_shape = get_batch_shape(x, self.event_shape)
        y2, ld, _ = prog.run_program(ops, x.reshape(-1, self.n_dim))
        return y2.reshape(x.shape), ld.reshape(batch_shape)

    # -- chunked evaluation (base.py:58-121) --------------------------------------------------------
    def batch_apply(self, fn: callable, batch_size: int, x: torch.Tensor, context: torch.Tensor = None, **kwargs):
        n_batch_dims = x.dim() - len(self.event_shape)
        xf = x.flatten(0, n_batch_dims - 1) if n_batch_dims > 1 else x
        cf = None
        if context is not None:
            cf = context.flatten(0, n_batch_dims - 1) if n_batch_dims > 1 else context
        outs: List[List[torch.Tensor]] = []
        for i in range(0, xf.shape[0], batch_size):
            args = (xf[i:i + batch_size],) if cf is None else (xf[i:i + batch_size], cf[i:i + batch_size])
            for j, piece in enumerate(fn(*args, **kwargs)):
                if len(outs) <= j:
                    outs.append([])
                outs[j].append(piece)
        return tuple(torch.cat(pieces, dim=0) for pieces in outs)

    def batch_forward(self, x: torch.Tensor, batch_size: int, context: torch.Tensor = None, **kwargs):
        return self.batch_apply(self.forward, batch_size, x, context, **kwargs)

    def batch_inverse(self, x: torch.Tensor, batch_size: int, context: torch.Tensor = None, **kwargs):
        return self.batch_apply(self.inverse, batch_size, x, context, **kwargs)

    def sq_norm_param(self) -> torch.Tensor:
        """Squared norm of the trainable parameters (base.py:134-144); on the GPU as one fused multi-tensor reduction
        forward and one multi-tensor scale backward instead of two tiny kernels per parameter tensor each way."""
        params = [p for p in self.parameters() if p.requires_grad]
        if params and all(p.is_cuda for p in params):
            return _SqNorm.apply(*params)
        return sum([torch.sum(torch.square(p)) for p in params])

    def regularization(self, *aux: Tuple[Any, ...]) -> torch.Tensor:
        return torch.tensor(0.0)

    def invert(self):
        self.forward, self.inverse = self.inverse, self.forward


class _SqNorm(torch.autograd.Function):
    """sum_i ||p_i||^2 over a list of CUDA tensors."""

    @staticmethod
    def forward(ctx, *params):
        ctx.save_for_backward(*params)
        norms = torch._foreach_norm([p.detach() for p in params], 2)
        return torch.stack(norms).square().sum()

    @staticmethod
    def backward(ctx, g):
        return tuple(torch._foreach_mul([p.detach() for p in ctx.saved_tensors], 2.0 * g))


def invert(bijection: Bijection) -> Bijection:
    bijection.forward, bijection.inverse = bijection.inverse, bijection.forward
    return bijection


class BijectiveComposition(Bijection):
    """Composition of bijections.  Consecutive lowerable layers are fused into one flow program."""

    def __init__(self, layers: List[Bijection], **kwargs):
        super().__init__(event_shape=layers[0].event_shape, context_shape=layers[0].context_shape)
        self.layers = nn.ModuleList(layers)

    def freeze_after(self, index: int):
        for i, layer in enumerate(self.layers):
            if i > index:
                layer.requires_grad_(False)

    def unfreeze_all_layers(self):
        for layer in self.layers:
            layer.requires_grad_(True)

    # -- fused execution --------------------------------------------------------------------------------
    def _segments(self, direction: str, context: torch.Tensor = None):
        """Split the layer sequence (in application order) into maximal runs of lowerable layers.
        A layer that needs a data-dependent initialisation first (ActNorm in training mode) starts a new run
        so that it can look at its own input.  ``context``: layers that can fold a context tensor into their lowered form
        (context-conditioned couplings: a per-row hidden bias) get it; the others ignore it or stay composite."""
        order = list(self.layers) if direction == 'forward' else list(self.layers)[::-1]
        segments, current = [], []
        for layer in order:
            if context is not None and getattr(layer, 'lowers_with_context', False):
                ops = layer.lower(direction, context=context)
            else:
                ops = layer.lower(direction)
            needs_data = direction == 'forward' and getattr(layer, 'needs_data_init', lambda: False)()
            # per-column layers the whole-flow kernels do not take (very wide events) still fuse with their neighbours
            # into one pass over the batch (csrc/b2f_colrun.cu)
            col = getattr(layer, 'column_op', lambda d: None)(direction) if ops is None else None
            if col is not None and not needs_data:
                if current:
                    segments.append(('ops', current))
                    current = []
                if segments and segments[-1][0] == 'cols':
                    segments[-1][1].append(col)
                else:
                    segments.append(('cols', [col]))
            elif ops is None or needs_data:
                if current:
                    segments.append(('ops', current))
                    current = []
                if ops is None and col is None and not needs_data:
                    segments.append(('layer', layer))
                elif ops is None:
                    segments.append(('layer_init', layer))   # initialised from its own input, then applied by itself
                else:
                    segments.append(('init', layer))   # lowered after its initialisation, see _run_layers
            else:
                current.extend(ops)
        if current:
            segments.append(('ops', current))
        return segments

    def fused_ops(self, direction: str) -> Optional[List[prog.LoweredOp]]:
        """The whole composition as one op list, or None if some layer cannot be lowered right now."""
        segs = self._segments(direction)
        if len(segs) == 1 and segs[0][0] == 'ops':
            return segs[0][1]
        if len(segs) == 0:
            return []
        return None

    def lower(self, direction: str):
        return self.fused_ops(direction)

    def _run_layers(self, x: torch.Tensor, context, direction: str, **kwargs):
        batch_shape = get_batch_shape(x, self.event_shape)
        x2 = x.reshape(-1, self.n_dim)
        log_det = None

        def add(ld):
            nonlocal log_det
            log_det = ld if log_det is None else log_det + ld

        for kind, item in self._segments(direction, context):
            if kind == 'ops':
                x2, ld, _ = prog.run_program(item, x2)
                add(ld)
            elif kind == 'cols':
                x2, ld = prog.run_column_ops(item, x2)
                add(ld)
            elif kind == 'layer_init':
                item.data_init(x2.reshape(*batch_shape, *self.event_shape),
                               reduce_fn=getattr(self, '_stats_reduce_fn', None))
                col = getattr(item, 'column_op', lambda d: None)(direction)
                if col is not None:
                    x2, ld = prog.run_column_ops([col], x2)
                    add(ld)
                else:
                    y, ld = item.forward(x2.reshape(*batch_shape, *self.event_shape), context=context, **kwargs)[:2]
                    x2 = y.reshape(-1, self.n_dim)
                    add(ld.reshape(-1))
            elif kind == 'init':
                item.data_init(x2.reshape(*batch_shape, *self.event_shape),
                               reduce_fn=getattr(self, '_stats_reduce_fn', None))
                x2, ld, _ = prog.run_program(item.lower(direction), x2)
                add(ld)
            else:
                fn = item.forward if direction == 'forward' else item.inverse
                y, ld = fn(x2.reshape(*batch_shape, *self.event_shape), context=context, **kwargs)[:2]
                x2 = y.reshape(-1, self.n_dim)
                add(ld.reshape(-1))
        if log_det is None:
            log_det = torch.zeros(x2.shape[0], device=x2.device, dtype=x2.dtype)
            x2 = x2.clone()
        return x2.reshape(x.shape), log_det.reshape(batch_shape)

    def forward(self, x: torch.Tensor, context: torch.Tensor = None, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._run_layers(x, context, "forward", **kwargs)

    def inverse(self, z: torch.Tensor, context: torch.Tensor = None, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._run_layers(z, context, "inverse")

    def regularization(self):
        """Sum of the layers' regularization terms (base.py:234-243).  Terms that live on the GPU are summed there; the
        layers' default `torch.tensor(0.0)` placeholders stay on the host (a 0-dim CPU tensor combines with a CUDA loss as
        a scalar, so no per-layer host-to-device copy is issued)."""
        terms = [t for t in (layer.regularization() for layer in self.layers) if isinstance(t, torch.Tensor)]
        gpu = [t for t in terms if t.is_cuda]
        cpu = [t for t in terms if not t.is_cuda]
        total = None
        for t in gpu:
            total = t if total is None else total + t
        if cpu:
            host = cpu[0]
            for t in cpu[1:]:
                host = host + t
            if total is None:
                return host
            if host.requires_grad or float(host) != 0.0:
                total = total + host.to(total.device)
        return total if total is not None else torch.tensor(0.0)
