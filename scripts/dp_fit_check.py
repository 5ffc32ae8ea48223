"""torchrun --nproc-per-node N scripts/dp_fit_check.py : data-parallel Flow.fit over NCCL must reproduce single-GPU
training with the global batch: same data, shuffle off, the loss of each of 20 steps within 1e-4 relative.
Ranks deliberately construct their flows under DIFFERENT seeds: fit() must broadcast rank 0's weights."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchflows_b200.flows as F  # noqa: E402
from torchflows_b200 import Flow  # noqa: E402
from torchflows_b200.architectures import CouplingRQNSF, RealNVP  # noqa: E402


def step_losses(flow, x, steps):
    """Loss before each of `steps` optimisation steps (full batch), through the same code path as fit()."""
    rank, world = F._dist_info()
    lo, hi = F.shard_bounds(len(x), rank, world)
    xs = x[lo:hi].to(flow.get_device())
    flow.train()
    flow._optimizer = torch.optim.AdamW(flow.parameters(), lr=0.01)
    return [float(flow.train_step(xs, n_global=len(x))) for _ in range(steps)]


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    ok = True
    for cls, D, kw in ((RealNVP, 16, {}), (CouplingRQNSF, 32, {}), (CouplingRQNSF, 64, {'conditioner_kwargs': {'n_hidden': 64}})):
        torch.manual_seed(0)
        x = torch.randn(4099, D) * 2 + 1
        torch.manual_seed(1)
        flow_1 = Flow(cls(D, **kw)).to(dev)
        sd0 = {k: v.clone() for k, v in flow_1.state_dict().items()}
        # (a) fit(): ranks start from different weights; fit broadcasts rank 0's and trains data-parallel
        torch.manual_seed(100 + rank)
        flow_dp = Flow(cls(D, **kw)).to(dev)
        if rank == 0:
            flow_dp.load_state_dict(sd0)
        flow_dp.fit(x, n_epochs=20, batch_size=None, shuffle=False, lr=0.01, keep_best_weights=False)
        with torch.no_grad():
            lp_dp = flow_dp.log_prob(x.to(dev)).mean().item()
        saved = F._dist_info
        F._dist_info = lambda: (0, 1)            # single-process reference run on every rank
        try:
            flow_1.fit(x, n_epochs=20, batch_size=None, shuffle=False, lr=0.01, keep_best_weights=False)
            with torch.no_grad():
                lp_1 = flow_1.log_prob(x.to(dev)).mean().item()
        finally:
            F._dist_info = saved
        # (b) per-step losses, DP vs single process, 20 steps from identical weights (ActNorm initialised by the first step)
        flow_a, flow_b = Flow(cls(D, **kw)).to(dev), Flow(cls(D, **kw)).to(dev)
        flow_a.load_state_dict(sd0)
        flow_b.load_state_dict(sd0)
        flow_a.bijection._stats_reduce_fn = F._allreduce_stats
        la = step_losses(flow_a, x, 20)
        flow_a.bijection._stats_reduce_fn = None
        F._dist_info = lambda: (0, 1)
        try:
            lb = step_losses(flow_b, x, 20)
        finally:
            F._dist_info = saved
        worst = max(abs(a - b) / (1 + abs(b)) for a, b in zip(la, lb))
        diff = torch.tensor([abs(lp_dp - lp_1) / (1 + abs(lp_1)), worst], device=dev)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        flat = torch.cat([p.detach().reshape(-1) for p in flow_dp.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        same_weights = bool(torch.equal(flat, ref))
        if rank == 0:
            print(f'{cls.__name__}({D}) world={world}: after fit mean log_prob DP {lp_dp:.6f} vs single {lp_1:.6f} '
                  f'(rel diff {diff[0].item():.2e}); worst per-step loss rel diff over 20 steps {diff[1].item():.2e}; '
                  f'weights identical on all ranks: {same_weights}', flush=True)
        ok &= diff[0].item() < 1e-3 and diff[1].item() < 1e-4 and same_weights
    if rank == 0:
        print('DP_FIT_OK' if ok else 'DP_FIT_MISMATCH', flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
