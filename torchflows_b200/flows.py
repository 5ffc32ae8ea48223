"""Distribution API: ``Flow(bijection).log_prob / sample / fit`` (API of torchflows/flows.py:18-455,606-713).

What changes underneath: ``log_prob`` and ``sample`` of a lowerable bijection are ONE kernel launch each
(layers + log-det + base density, csrc/b2f_flow.cu); ``fit`` trains through the hand-written backward kernel
and, when ``torch.distributed`` is initialised, runs data-parallel: every rank takes its slice of each
minibatch and gradients are all-reduced over NCCL (NVLink/NVSwitch) in one flat bucket per step."""
import math
import time
import warnings
from typing import Tuple, Union

import torch
import torch.nn as nn
from tqdm import tqdm

from torchflows_b200 import _native as N
from torchflows_b200 import _program as prog
from torchflows_b200.base_distributions.gaussian import DiagonalGaussian
from torchflows_b200.bijections.base import Bijection
from torchflows_b200.utils import event_size, flatten_event, get_batch_shape, unflatten_event


def _dist_info() -> Tuple[int, int]:
    """(rank, world_size) of the default process group, (0, 1) when not distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of an n-row batch owned by `rank`; sizes differ by at most one row."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients(parameters, world: int) -> None:
    """Mean of the gradients over all ranks in one flat bucket (kept for callers that hold plain ``.grad`` tensors; ``fit``
    and ``train_step`` use ``GradBuckets``, which overlaps the exchange with the backward pass and copies nothing)."""
    import torch.distributed as dist
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads or world == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    offset = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[offset:offset + n].view_as(g))
        offset += n


class _GaussLogProb(torch.autograd.Function):
    """DiagonalGaussian.log_prob over rows with fixed parameters (gaussian.py:46-54): b2f_gauss_log_prob / _backward."""

    @staticmethod
    def forward(ctx, z2, loc, log_scale):
        ctx.save_for_backward(z2, loc, log_scale)
        return N.gauss_log_prob(z2, loc, log_scale)

    @staticmethod
    def backward(ctx, g):
        z2, loc, log_scale = ctx.saved_tensors
        return N.gauss_log_prob_backward(z2, loc, log_scale, g.contiguous()), None, None


class GradBuckets:
    """Data-parallel gradient exchange of ``Flow.fit``: one bucket per layer of the bijection, gradients
    LIVING in the bucket (every ``p.grad`` is a view into its bucket's flat buffer, so nothing is packed or copied back),
    and each bucket's all-reduce launched from a post-accumulate hook the moment the backward pass has produced the last
    gradient of that layer -- it runs on NCCL's stream while autograd is still working on the layers in front of it.
    ``finish()`` waits for the exchanges before the optimizer step.  Mean over ranks: ``ReduceOp.AVG`` on NCCL (no extra
    pass), sum + scale elsewhere (gloo, CPU tests)."""

    MIN_BUCKET_BYTES = 1 << 20      # neighbouring layers are merged up to this size: launch latency, not link count

    def __init__(self, flow, world: int):
        import torch.distributed as dist
        self.world = world
        self.avg = dist.get_backend() == 'nccl'
        groups, seen = [], set()
        layers = list(getattr(flow.bijection, 'layers', [flow.bijection]))
        for layer in reversed(layers):           # backward finishes the last layer first
            ps = [p for p in layer.parameters() if p.requires_grad and id(p) not in seen]
            seen.update(id(p) for p in ps)
            if ps:
                groups.append(ps)
        rest = [p for p in flow.parameters() if p.requires_grad and id(p) not in seen]
        if rest:
            groups.append(rest)
        merged = []
        for ps in groups:
            size = sum(p.numel() * p.element_size() for p in ps)
            if merged and merged[-1][1] < self.MIN_BUCKET_BYTES:
                merged[-1][0].extend(ps)
                merged[-1][1] += size
            else:
                merged.append([list(ps), size])
        self.flats, self.params, self.pending, self.bucket_of, self.hooks, self.handles = [], [], [], {}, [], []
        for bi, (ps, _) in enumerate(merged):
            flat = torch.zeros(sum(p.numel() for p in ps), device=ps[0].device, dtype=ps[0].dtype)
            offset = 0
            for p in ps:
                p.grad = flat[offset:offset + p.numel()].view_as(p)
                offset += p.numel()
                self.bucket_of[id(p)] = bi
                self.hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
            self.flats.append(flat)
            self.params.append(ps)
            self.pending.append(len(ps))
        self.launched = [False] * len(self.flats)
        self.nbytes = sum(f.numel() * f.element_size() for f in self.flats)

    def begin_step(self):
        """Instead of optimizer.zero_grad(): the gradients stay views of the buckets."""
        torch._foreach_zero_(self.flats)
        self.pending = [len(ps) for ps in self.params]
        self.launched = [False] * len(self.flats)
        self.handles = []

    def _launch(self, bi):
        import torch.distributed as dist
        flat = self.flats[bi]
        offset = 0
        for p in self.params[bi]:               # a gradient that autograd re-bound instead of accumulating in place
            view = flat[offset:offset + p.numel()]
            if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad.reshape(-1))
                p.grad = view.view_as(p)
            offset += p.numel()
        self.handles.append(dist.all_reduce(flat, op=dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM, async_op=True))
        self.launched[bi] = True

    def _on_grad(self, p):
        bi = self.bucket_of[id(p)]
        self.pending[bi] -= 1
        if self.pending[bi] == 0 and not self.launched[bi]:
            self._launch(bi)

    def finish(self):
        for bi in range(len(self.flats)):       # layers that received no gradient this step still take part (zeros)
            if not self.launched[bi]:
                self._launch(bi)
        for h in self.handles:
            h.wait()
        if not self.avg:
            torch._foreach_div_(self.flats, float(self.world))
        self.handles = []

    def close(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []


def _allreduce_stats(s, q, n):
    import torch.distributed as dist
    packed = torch.cat([s, q, n.reshape(1)])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    d = s.numel()
    return packed[:d], packed[d:2 * d], packed[2 * d]


def _make_adamw(params, lr, capturable=False):
    """The reference's optimizer: `torch.optim.AdamW(self.parameters(), lr=lr)` (flows.py:309,537).  The fused
    single-kernel variant is deliberately NOT used: with it stochastic variational inference on the affine presets stopped
    converging in tests/test_gpu_training.py::test_variational_fit_and_kl_fit (0 of 6 initialisations against 6 of 6).
    `capturable`: step counter on the device, so that optimizer.step() can be part of a CUDA graph."""
    return torch.optim.AdamW(list(params), lr=lr, capturable=capturable)


class _Snapshot(dict):
    """A copy of a state_dict that remembers the live tensors it was taken from."""
    sources = None


def _touch(params):
    """Bump the version counters of parameters that were updated inside a CUDA-graph replay (no Python ran): the kernel
    operand layouts cached per (pointer, version) in _program.py are derived again on next use."""
    params = [p for p in params if p.numel() > 0]
    if params:
        with torch.no_grad():
            torch._foreach_add_(params, 0.0)


class _GraphTrainStep:
    """One training step of `BaseFlow.fit` -- loss, backward, AdamW -- captured in a CUDA graph for a fixed batch size.
    The kernels of this package launch on the current stream through the C ABI and allocate through torch, so they are
    captured like any torch op; operand layouts derived from the weights are recomputed inside the graph."""

    def __init__(self, flow, xb, wb):
        self.rows = len(xb)
        self.x, self.w = xb.clone(), wb.clone()
        opt = flow._optimizer
        opt.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = flow._base_batch_loss((self.x, self.w), reduction=torch.mean, use_regularization=True)
            self.loss.backward()
            opt.step()

    def __call__(self, xb, wb):
        self.x.copy_(xb)
        self.w.copy_(wb)
        self.graph.replay()
        return self.loss.detach()


class BaseFlow(nn.Module):
    def __init__(self, event_shape, base_distribution: Union[torch.distributions.Distribution, str] = 'standard_normal'):
        super().__init__()
        self.event_shape = event_shape
        self.event_size = event_size(event_shape)
        if isinstance(base_distribution, str) and base_distribution == 'standard_normal':
            self.base = DiagonalGaussian(loc=torch.zeros(self.event_size), scale=torch.ones(self.event_size))
        elif isinstance(base_distribution, torch.distributions.Distribution):
            self.base = base_distribution
        else:
            raise ValueError(f'Invalid base distribution: {base_distribution}')
        self.register_buffer('device_buffer', torch.empty(size=()))
        self._optimizer = None
        self._buckets = None        # GradBuckets while training data-parallel

    def get_device(self):
        return self.device_buffer.device

    def base_log_prob(self, z: torch.Tensor):
        zf = flatten_event(z, self.event_shape)
        if zf.is_cuda and zf.dtype == torch.float32 and self._fusable_base():
            # one kernel each way instead of six elementwise passes (flows whose layers are not a single program)
            batch_shape = zf.shape[:-1]
            lp = _GaussLogProb.apply(zf.reshape(-1, zf.shape[-1]), self.base.loc.detach().reshape(-1).contiguous(),
                                     self.base.log_scale.detach().reshape(-1).contiguous())
            return lp.reshape(batch_shape)
        return self.base.log_prob(zf)

    def base_sample(self, sample_shape: Union[torch.Size, Tuple[int, ...]]):
        return unflatten_event(self.base.sample(sample_shape), self.event_shape)

    def regularization(self, *args, **kwargs):
        return self.bijection.regularization(*args, **kwargs)

    def _fusable_base(self) -> bool:
        b = self.base
        return type(b) is DiagonalGaussian and not any(p.requires_grad for p in b.parameters())

    # ---------------------------------------------------------------------------------------------------
    def _base_batch_loss(self, batch, reduction: callable = torch.mean, use_regularization: bool = True):
        """-reduction(w * log_prob(x)) / event_size + regularization   (flows.py:199-224)."""
        x, weights = batch[:2]
        context = batch[2] if len(batch) == 3 else None
        log_prob = self.log_prob(x.to(self.get_device()), context=context)
        loss = -reduction(log_prob * weights.to(self.get_device())) / self.event_size
        if use_regularization:
            loss = loss + self.regularization()
        return loss

    def _snapshot_weights(self, into: dict = None) -> dict:
        """`deepcopy(self.state_dict())` of the reference's keep-best-weights logic (flows.py:246,429) without the
        per-tensor Python work: the first call clones, later calls refresh the same buffers with one fused copy."""
        if into is None or getattr(into, 'sources', None) is None:
            state = self.state_dict()
            snap = _Snapshot((k, v.detach().clone()) for k, v in state.items())
            snap.sources = [v.detach() for v in state.values()]     # the module's own parameters and buffers
            return snap
        dst = list(into.values())
        if dst:
            torch._foreach_copy_(dst, into.sources)
        return into

    def fit(self, x_train: torch.Tensor, n_epochs: int = 500, lr: float = 0.05, batch_size: Union[int, str] = 1024,
            shuffle: bool = True, show_progress: bool = False, w_train: torch.Tensor = None,
            context_train: torch.Tensor = None, x_val: torch.Tensor = None, w_val: torch.Tensor = None,
            context_val: torch.Tensor = None, keep_best_weights: bool = True, early_stopping: bool = False,
            early_stopping_threshold: int = 50, max_batch_size_mb: int = None,
            time_limit_seconds: Union[float, int] = None, reset_optimizer: bool = True, cuda_graph: bool = False):
        """See ``_fit`` for the training loop.  This wrapper owns what data-parallel training adds around it and undoes it
        whatever happens inside: every rank starts from rank 0's weights and buffers and shuffles with rank 0's seed,
        ActNorm initialises from all-reduced statistics, gradients are exchanged through ``GradBuckets``."""
        rank, world = _dist_info()
        seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64)      # from the global RNG, like a DataLoader's shuffle
        trains = any(p.requires_grad for p in self.parameters())
        if world > 1 and trains:
            import torch.distributed as dist
            device = self.get_device()
            for t in self.state_dict().values():
                if t.numel() > 0:
                    dist.broadcast(t, src=0)
            seed_dev = seed.to(device)
            dist.broadcast(seed_dev, src=0)
            seed = seed_dev.cpu()
            self.bijection._stats_reduce_fn = _allreduce_stats
            self._buckets = GradBuckets(self, world)
        try:
            return self._fit(x_train, n_epochs, lr, batch_size, shuffle, show_progress, w_train, context_train, x_val, w_val,
                             context_val, keep_best_weights, early_stopping, early_stopping_threshold, max_batch_size_mb,
                             time_limit_seconds, reset_optimizer, cuda_graph, int(seed))
        finally:
            # a later ActNorm initialisation or train_step outside this process group must not try to communicate
            if world > 1 and trains:
                self.bijection._stats_reduce_fn = None
                if self._buckets is not None:
                    self._buckets.close()
                self._buckets = None

    def _fit(self, x_train, n_epochs, lr, batch_size, shuffle, show_progress, w_train, context_train, x_val, w_val,
             context_val, keep_best_weights, early_stopping, early_stopping_threshold, max_batch_size_mb,
             time_limit_seconds, reset_optimizer, cuda_graph, shuffle_seed):
        """Maximum-likelihood fit with the reference's semantics (flows.py:226-455): AdamW, minibatches in a fresh
        random order every epoch, best-weights snapshot, divergence rollback, optional validation / early stopping /
        adaptive batch size / time limit.  The data are moved to the flow's device once and batches are index
        slices there (the reference collates every batch on the host through a DataLoader).

        ``cuda_graph=True`` (not in the reference; single process, no context): after three ordinary steps the training
        step (forward, backward, AdamW) is captured once per batch size in a CUDA graph and replayed, so a step costs
        one graph launch instead of ~60 kernel launches and their Python -- for launch-bound problems such as the README
        example.  Same arithmetic; the loss is checked for divergence after the replayed step instead of before its
        backward pass (a diverged step is rolled back to the best weights either way)."""
        t0 = time.time()
        self.train()
        params = list(self.parameters())
        if len(params) == 0:
            return
        if not any(p.requires_grad for p in params):
            self.eval()
            return
        device = self.get_device()
        rank, world = _dist_info()
        n_train = len(x_train)

        adaptive = isinstance(batch_size, str) and batch_size == 'adaptive'
        if batch_size is None:
            batch_size = n_train
        elif adaptive:
            max_batch_size = min(4096, n_train // 10)
            if max_batch_size_mb is not None:
                max_batch_size = max(1, min(max_batch_size, int(max_batch_size_mb / (self.event_size / 2 ** 20))))
            batch_size = max(32, min(1024, n_train // 100))

        x_dev = x_train.to(device=device, dtype=torch.float32)
        w_dev = (torch.ones(n_train) if w_train is None else w_train).to(device=device, dtype=torch.float32)
        if len(w_dev) != n_train:
            raise ValueError(f'Expected same number of training data and training weights, '
                             f'but found {n_train} and {len(w_dev)}')
        c_dev = None
        if context_train is not None:
            if len(context_train) != n_train:
                raise ValueError(f'Expected same number of training data and training contexts, '
                                 f'but found {n_train} and {len(context_train)}')
            c_dev = context_train.to(device=device, dtype=torch.float32)
        cv_dev = None
        if x_val is not None and context_val is not None:
            if len(context_val) != len(x_val):
                raise ValueError(f'Expected same number of validation data and validation contexts, '
                                 f'but found {len(x_val)} and {len(context_val)}')
            cv_dev = context_val.to(device=device, dtype=torch.float32)
        if x_val is not None:
            xv_dev = x_val.to(device=device, dtype=torch.float32)
            wv_dev = (torch.ones(len(x_val)) if w_val is None else w_val).to(device=device, dtype=torch.float32)
            if len(wv_dev) != len(xv_dev):
                raise ValueError(f'Expected same number of validation data and validation weights, '
                                 f'but found {len(xv_dev)} and {len(wv_dev)}')

        use_graph = bool(cuda_graph) and world == 1 and context_train is None and device.type == 'cuda'
        if self._optimizer is None or reset_optimizer:
            self._optimizer = _make_adamw(self.parameters(), lr, capturable=use_graph)
        elif use_graph and not all(g.get('capturable', False) for g in self._optimizer.param_groups):
            use_graph = False        # an optimizer kept from an earlier call cannot be stepped inside a graph: stay eager
        trainable = [p for p in self.parameters() if p.requires_grad]
        graph_step, eager_steps, replayed, loss = None, 0, False, None
        buckets = self._buckets
        gen = torch.Generator(device='cpu')
        gen.manual_seed(shuffle_seed)                               # identical batch order on every rank

        val_loss = None
        best_val_loss = best_train_loss = float('inf')
        best_val_epoch = best_train_epoch = 0
        best_weights = self._snapshot_weights()
        diverged = False

        for epoch in (pbar := tqdm(range(n_epochs), desc='Fitting NF', disable=not show_progress)):
            out_of_time = time_limit_seconds is not None and time.time() - t0 >= time_limit_seconds
            if world > 1 and time_limit_seconds is not None:
                import torch.distributed as dist
                flag = torch.tensor([float(out_of_time)], device=device)      # ranks stop together or not at all: one rank
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)                  # leaving alone would hang the others' collectives
                out_of_time = bool(flag.item())
            if out_of_time:
                print('Training time limit exceeded')
                break
            if adaptive and epoch % 10 == 9 and batch_size < max_batch_size:
                batch_size = min(batch_size * 2, max_batch_size)      # doubled every 10 epochs (flows.py:346-352)

            order = torch.randperm(n_train, generator=gen).to(device) if shuffle else None
            total, n_batches = 0.0, 0
            for start in range(0, n_train, batch_size):
                stop = min(start + batch_size, n_train)
                lo, hi = shard_bounds(stop - start, rank, world)
                idx = slice(start + lo, start + hi) if order is None else order[start + lo:start + hi]
                xb, wb = x_dev[idx], w_dev[idx]
                cb = None if c_dev is None else c_dev[idx]
                if use_graph and eager_steps >= 3 and stop - start == min(batch_size, n_train):
                    if graph_step is None or graph_step.rows != len(xb):
                        # operand layouts cached for the current weights (e.g. by the validation pass) must be derived
                        # again INSIDE the capture, or every replay would read the frozen copies
                        _touch(trainable)
                        # no reference to an eager step's autograd graph may survive: its AccumulateGrad nodes are bound
                        # to the stream they were created on and would break the capture
                        loss = graph_step = None
                        graph_step = _GraphTrainStep(self, xb, wb)
                    loss_value = graph_step(xb, wb).item()        # one host sync per step
                    replayed = True
                    if not math.isfinite(loss_value):
                        _touch(trainable)
                        self.load_state_dict(best_weights)
                        warnings.warn('Flow training diverged. Reverting to previous weights.')
                        diverged = True
                        break
                    total += loss_value
                    n_batches += 1
                    continue
                if replayed:
                    _touch(trainable)          # the replays updated the weights behind the operand caches' back
                    replayed = False
                eager_steps += 1
                loss, loss_value = self._loss_and_backward_inputs(xb, wb, stop - start, world, cb)
                if not torch.isfinite(loss_value):
                    self.load_state_dict(best_weights)       # roll back (flows.py:387-393)
                    warnings.warn('Flow training diverged. Reverting to previous weights.')
                    diverged = True
                    break
                total += float(loss_value)
                n_batches += 1
                loss.backward()
                if buckets is not None:
                    buckets.finish()
                self._optimizer.step()
                if show_progress:
                    msg = f'Training loss (batch): {float(loss_value):.4f} [{best_train_loss:.4f} @ {best_train_epoch}]'
                    if val_loss is not None:
                        msg += f' , Validation loss (batch): {val_loss:.4f} [{best_val_loss:.4f} @ {best_val_epoch}]'
                    pbar.set_postfix_str(msg)
            if diverged:
                break

            mean_loss = total / max(n_batches, 1)
            if mean_loss < best_train_loss:
                best_train_loss, best_train_epoch = mean_loss, epoch
            if replayed and x_val is not None:
                _touch(trainable)
                replayed = False
            if x_val is not None:
                with torch.no_grad():
                    acc = 0.0
                    for start in range(0, len(xv_dev), batch_size):
                        sl = slice(start, start + batch_size)
                        vb = (xv_dev[sl], wv_dev[sl]) if cv_dev is None else (xv_dev[sl], wv_dev[sl], cv_dev[sl])
                        acc += float(self._base_batch_loss(vb, reduction=torch.sum, use_regularization=False))
                val_loss = acc / len(xv_dev)
                if val_loss < best_val_loss:
                    best_val_loss, best_val_epoch = val_loss, epoch
            if keep_best_weights:
                improved = best_val_epoch == epoch if x_val is not None else best_train_epoch == epoch
                if improved:
                    best_weights = self._snapshot_weights(best_weights)
            if early_stopping:
                ref_epoch = best_val_epoch if x_val is not None else best_train_epoch
                if epoch - ref_epoch > early_stopping_threshold:
                    break

        if replayed:
            _touch(trainable)
        graph_step = None
        if keep_best_weights:
            self.load_state_dict(best_weights)
        self.eval()

    def _loss_and_backward_inputs(self, xb, wb, n_global: int, world: int, cb=None):
        """Loss of this rank's slice of a minibatch, scaled so that the mean over ranks of the gradients is the gradient
        of the global-batch loss -mean(w*log_prob)/event_size + regularization (flows.py:199-224); returns the local
        loss tensor (to call backward on) and the global loss value (all-reduced when world > 1)."""
        buckets = getattr(self, '_buckets', None)
        if buckets is not None:
            buckets.begin_step()                 # gradients are views of the all-reduce buckets: zero them in place
        else:
            self._optimizer.zero_grad()
        if world == 1:
            batch = (xb, wb) if cb is None else (xb, wb, cb)
            loss = self._base_batch_loss(batch, reduction=torch.mean, use_regularization=True)
            return loss, loss.detach()
        import torch.distributed as dist
        lp = self.log_prob(xb, context=cb)
        loss = -(lp * wb).sum() * (world / n_global) / self.event_size + self.regularization()
        loss_value = loss.detach().clone()
        dist.all_reduce(loss_value, op=dist.ReduceOp.SUM)
        loss_value /= world
        return loss, loss_value

    def train_step(self, xb: torch.Tensor, wb: torch.Tensor = None, n_global: int = None):
        """One optimisation step exactly as the inner loop of ``fit`` does it (forward, backward, gradient all-reduce
        when distributed, AdamW).  ``xb`` is this rank's slice; ``n_global`` the size of the global minibatch."""
        rank, world = _dist_info()
        if self._optimizer is None:
            self._optimizer = _make_adamw(self.parameters(), 0.05)
        if wb is None:
            wb = torch.ones(len(xb), device=xb.device)
        if world > 1 and getattr(self, '_buckets', None) is None:
            self._buckets = GradBuckets(self, world)           # kept for the following steps; fit() builds its own
        loss, loss_value = self._loss_and_backward_inputs(xb, wb, n_global or len(xb) * world, world)
        loss.backward()
        if world > 1:
            self._buckets.finish()
        self._optimizer.step()
        return loss_value

    # -- KL(p || q) and stochastic variational inference (flows.py:81-197, 457-603) --------------------------------
    def _loss_kl_p_to_q(self, data: torch.Tensor, log_prob_target_data: torch.Tensor, use_regularization: bool = True):
        """mean(log p(x) - log q(x)) + regularization   (flows.py:81-94)."""
        dev = self.get_device()
        loss = torch.mean(log_prob_target_data.to(dev) - self.log_prob(data.to(dev)))
        if use_regularization:
            loss = loss + self.regularization()
        return loss

    def fit_kl_p_to_q(self, x_train: torch.Tensor, x_val: torch.Tensor, neg_log_prob_target: callable,
                      n_epochs: int = 500, lr: float = 0.05, batch_size: int = 1024, show_progress: bool = False,
                      keep_best_weights: bool = True, early_stopping: bool = False, early_stopping_threshold: int = 50,
                      time_limit_seconds: float = None, reset_optimizer: bool = True):
        """Fit by minimising KL(p || q) on samples of p with known (negative) log density (flows.py:96-197): batches in
        the given order, validation loss summed over the validation batches, best weights by validation loss."""
        lp_train = -neg_log_prob_target(x_train).detach()
        lp_val = -neg_log_prob_target(x_val).detach()
        if len(list(self.parameters())) == 0:
            return
        self.train()
        t0 = time.time()
        if self._optimizer is None or reset_optimizer:
            self._optimizer = _make_adamw(self.parameters(), lr)
        val_loss, best_val_loss, best_epoch = None, float('inf'), 0
        best_weights = self._snapshot_weights()
        for epoch in (pbar := tqdm(range(n_epochs), desc='Fitting NF', disable=not show_progress)):
            if time_limit_seconds is not None and time.time() - t0 >= time_limit_seconds:
                print('Training time limit exceeded')
                break
            for s0 in range(0, len(x_train), batch_size):
                self._optimizer.zero_grad()
                loss = self._loss_kl_p_to_q(x_train[s0:s0 + batch_size], lp_train[s0:s0 + batch_size])
                loss.backward()
                self._optimizer.step()
                if show_progress:
                    pbar.set_postfix_str(f'Training loss (batch): {float(loss):.4f}')
            with torch.no_grad():
                val_loss = 0.0
                for s0 in range(0, len(x_val), batch_size):
                    val_loss += float(self._loss_kl_p_to_q(x_val[s0:s0 + batch_size], lp_val[s0:s0 + batch_size],
                                                           use_regularization=False))
            if val_loss < best_val_loss:
                best_val_loss, best_epoch = val_loss, epoch
            if keep_best_weights and best_epoch == epoch:
                best_weights = self._snapshot_weights(best_weights)
            if early_stopping and epoch - best_epoch > early_stopping_threshold:
                break
        if keep_best_weights:
            self.load_state_dict(best_weights)
        self.eval()

    def _variational_loss(self, target_log_prob: callable, n_samples: int, use_regularization: bool = True,
                          check_for_divergences: bool = False):
        """-mean(log p(x) + flow_log_prob) over x ~ flow, where flow_log_prob is what sample(return_log_prob=True)
        returns (flows.py:457-493, incl. the reference's sign convention)."""
        flow_x, flow_log_prob = self.sample(n_samples, return_log_prob=True)
        target_value = target_log_prob(flow_x)
        loss = -torch.mean(target_value + flow_log_prob)
        if use_regularization:
            loss = loss + self.regularization()
        diverged = False
        if check_for_divergences:
            diverged = bool((~torch.isfinite(loss)) | (flow_x.abs().max() > 1e8) | (flow_log_prob.abs().max() > 1e6)
                            | (~torch.isfinite(flow_x)).any() | (~torch.isfinite(flow_log_prob)).any())
        return loss, flow_log_prob, target_value, diverged

    def variational_fit(self, target_log_prob: callable, n_epochs: int = 500, lr: float = 0.05, n_samples: int = 1,
                        early_stopping: bool = False, early_stopping_threshold: int = 50,
                        keep_best_weights: bool = True, show_progress: bool = False,
                        check_for_divergences: bool = False, time_limit_seconds: Union[float, int] = None,
                        reset_optimizer: bool = True):
        """Stochastic variational inference (flows.py:495-603): one reparameterised sample batch per epoch, divergent
        epochs are skipped, non-finite parameters revert to the initial weights, best weights by training loss.
        Gradients flow through the sampling direction: coupling layers (incl. the inverse spline) and
        one-pass MADE layers (IAF) are fused; the sequential MADE direction (MAF sampling) has no fused backward."""
        t0 = time.time()
        if len(list(self.parameters())) == 0:
            return
        self.train()
        if self._optimizer is None or reset_optimizer:
            self._optimizer = _make_adamw(self.parameters(), lr)
        best_loss, best_epoch, n_divergences, reverted = float('inf'), 0, 0, False
        initial_weights = self._snapshot_weights()
        best_weights = self._snapshot_weights()
        for epoch in (pbar := tqdm(range(n_epochs), desc='Fitting with SVI', disable=not show_progress)):
            if time_limit_seconds is not None and time.time() - t0 >= time_limit_seconds:
                print('Training time limit exceeded')
                break
            if check_for_divergences and not all(bool(torch.isfinite(p).all()) for p in self.parameters()):
                print('Flow training diverged')
                print('Reverting to initial weights')
                reverted = True
                break
            self._optimizer.zero_grad()
            loss_value = float('nan')
            try:
                loss, flow_lp, target_lp, diverged = self._variational_loss(target_log_prob, n_samples, True, True)
                if not diverged:
                    loss.backward()
                    self._optimizer.step()
                    loss_value = float(loss.detach())
                    if loss_value < best_loss:
                        best_loss, best_epoch = loss_value, epoch
                        if keep_best_weights:
                            best_weights = self._snapshot_weights(best_weights)
            except ValueError:
                diverged = True
            n_divergences += int(diverged)
            if show_progress:
                pbar.set_postfix_str(f'Loss: {loss_value:.4f} [best: {best_loss:.4f} @ {best_epoch}], '
                                     f'divergences: {n_divergences}')
            if early_stopping and epoch - best_epoch > early_stopping_threshold:
                break
        if reverted:
            self.load_state_dict(initial_weights)
        elif keep_best_weights:
            self.load_state_dict(best_weights)
        self.eval()


class Flow(BaseFlow):
    """A bijection applied to a base distribution."""

    def __init__(self, bijection: Bijection, **kwargs):
        super().__init__(event_shape=bijection.event_shape, **kwargs)
        self.register_module('bijection', bijection)

    @property
    def context_shape(self):
        return self.bijection.context_shape

    def forward_with_log_prob(self, x: torch.Tensor, context: torch.Tensor = None):
        """z = bijection.forward(x); log_prob = base.log_prob(z) + log_det   (flows.py:628-648)."""
        if context is not None:
            if self.context_shape is None:
                raise ValueError('Context shape must be set.')
            assert get_batch_shape(x, self.event_shape) == get_batch_shape(context, self.context_shape)
            context = context.to(self.get_device())
        x = x.to(self.get_device())
        ops = self.bijection.lower('forward') if context is None else None
        if ops is not None and self._fusable_base():
            batch_shape = get_batch_shape(x, self.event_shape)
            z2, _, lp = prog.run_program(ops, x.reshape(-1, self.event_size), want_log_prob=True,
                                         base_loc=self.base.loc, base_log_scale=self.base.log_scale)
            return z2.reshape(x.shape), lp.reshape(batch_shape)
        z, log_det = self.bijection.forward(x, context=context)[:2]
        return z, self.base_log_prob(z) + log_det

    def log_prob(self, x: torch.Tensor, context: torch.Tensor = None) -> torch.Tensor:
        if context is None and not (torch.is_grad_enabled() and (x.requires_grad or any(
                p.requires_grad for p in self.parameters()))):
            # inference: skip writing z back to HBM altogether
            x = x.to(self.get_device())
            ops = self.bijection.lower('forward')
            if ops is not None and self._fusable_base() and len(ops) <= N.MAX_OPS:
                batch_shape = get_batch_shape(x, self.event_shape)
                _, _, lp = prog.run_program(ops, x.reshape(-1, self.event_size), want_log_prob=True,
                                            base_loc=self.base.loc, base_log_scale=self.base.log_scale, want_y=False)
                return lp.reshape(batch_shape)
        return self.forward_with_log_prob(x, context)[1]

    def _base_draw_parameters(self):
        """(loc, log_scale) for the library's base draws, or (None, None) for a standard normal base (both all zero): the
        kernels then skip the per-draw scale and shift.  One device read per parameter version, cached."""
        loc, log_scale = self.base.loc, self.base.log_scale
        key = (loc.data_ptr(), loc._version, log_scale.data_ptr(), log_scale._version)
        hit = getattr(self, '_b2f_std_base', None)
        if hit is None or hit[0] != key:
            hit = (key, not bool(loc.detach().any()) and not bool(log_scale.detach().any()))
            object.__setattr__(self, '_b2f_std_base', hit)
        return (None, None) if hit[1] else (loc, log_scale)

    def sample(self, sample_shape: Union[int, torch.Size, Tuple[int, ...]], context: torch.Tensor = None,
               no_grad: bool = False, return_log_prob: bool = False):
        """x = bijection.inverse(z), z ~ base.  With ``return_log_prob`` the second output is
        ``base.log_prob(z) + log_det_inverse`` exactly as in the reference (flows.py:710-712)."""
        if isinstance(sample_shape, int):
            sample_shape = (sample_shape,)
        sample_shape = tuple(sample_shape)
        if context is None:
            grad_needed = (not no_grad) and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
            dev = self.get_device()
            if not grad_needed and dev.type == 'cuda' and self._fusable_base():
                # inference: the base draws are made by the library (counter-based Philox stream keyed by torch's seed);
                # spline coupling programs draw them inside the kernel, the noise never touches memory
                ops = self.bijection.lower('inverse')
                if ops is not None and len(ops) <= N.MAX_OPS:
                    n = 1
                    for d in sample_shape:
                        n *= int(d)
                    loc, log_scale = self._base_draw_parameters()
                    with torch.no_grad():
                        x2, lp = prog.run_sample_program(ops, n, self.event_size, dev, want_log_prob=return_log_prob,
                                                         base_loc=loc, base_log_scale=log_scale)
                    x = x2.reshape(*sample_shape, *self.event_shape)
                    return (x, lp.reshape(sample_shape)) if return_log_prob else x
            z = self.base_sample(sample_shape=sample_shape)
            return self._sample_from_base(z, no_grad, return_log_prob)
        # context-conditioned (flows.py:680-692): either one context per sampled element, or one context tensor per
        # entry of `context` for which `sample_shape` samples each are drawn
        context = context.to(self.get_device())
        if tuple(get_batch_shape(context, self.context_shape)) != sample_shape:
            z = self.base_sample(sample_shape=(*sample_shape, len(context)))
            context = context[None].expand(*sample_shape, *context.shape).contiguous() if len(sample_shape) == 1 else \
                context[(None,) * len(sample_shape)].expand(*sample_shape, *context.shape).contiguous()
        else:
            z = self.base_sample(sample_shape=sample_shape)
        if no_grad:
            with torch.no_grad():
                x, log_det = self.bijection.inverse(z.detach(), context=context)[:2]
        else:
            x, log_det = self.bijection.inverse(z, context=context)[:2]
        if return_log_prob:
            return x, self.base_log_prob(z) + log_det
        return x

    def _sample_from_base(self, z: torch.Tensor, no_grad: bool = False, return_log_prob: bool = False):
        """The deterministic part of ``sample``: push base draws ``z`` through the inverse bijection."""
        z = z.to(self.get_device())
        batch_shape = get_batch_shape(z, self.event_shape)
        grad_needed = (not no_grad) and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if not grad_needed:
            ops = self.bijection.lower('inverse')
            if ops is not None and self._fusable_base() and len(ops) <= N.MAX_OPS:
                with torch.no_grad():
                    x2, _, lp = prog.run_program(ops, z.detach().reshape(-1, self.event_size),
                                                 want_log_prob=return_log_prob, base_loc=self.base.loc,
                                                 base_log_scale=self.base.log_scale, flags=N.FLOW_LOGP_OF_INPUT)
                x = x2.reshape(z.shape)
                return (x, lp.reshape(batch_shape)) if return_log_prob else x
        if no_grad:
            with torch.no_grad():
                x, log_det = self.bijection.inverse(z.detach())[:2]
        else:
            x, log_det = self.bijection.inverse(z)[:2]
        if return_log_prob:
            return x, self.base_log_prob(z) + log_det
        return x
