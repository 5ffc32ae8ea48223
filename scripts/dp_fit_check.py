"""torchrun --nproc-per-node N scripts/dp_fit_check.py : data-parallel Flow.fit over NCCL must reproduce single-GPU
training with the global batch: same data, same init, shuffle off, 20 steps."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchflows_b200 import Flow  # noqa: E402
from torchflows_b200.architectures import CouplingRQNSF, RealNVP  # noqa: E402


def losses_after_fit(flow, x, steps):
    flow.fit(x, n_epochs=steps, batch_size=None, shuffle=False, lr=0.01, keep_best_weights=False)
    with torch.no_grad():
        return flow.log_prob(x.to(flow.get_device())).mean().item()


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    ok = True
    for cls, D in ((RealNVP, 16), (CouplingRQNSF, 32)):
        torch.manual_seed(0)
        x = torch.randn(4099, D) * 2 + 1
        torch.manual_seed(1)
        flow_dp = Flow(cls(D)).to(dev)
        lp_dp = losses_after_fit(flow_dp, x, 20)
        # single-process reference run on every rank (no process group visible to fit)
        torch.manual_seed(1)
        flow_1 = Flow(cls(D)).to(dev)
        import torchflows_b200.flows as F
        saved = F._dist_info
        F._dist_info = lambda: (0, 1)
        lp_1 = losses_after_fit(flow_1, x, 20)
        F._dist_info = saved
        same = torch.tensor([abs(lp_dp - lp_1)], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f'{cls.__name__}({D}) world={world}: mean log_prob DP {lp_dp:.6f} vs single {lp_1:.6f} (max diff over ranks {same.item():.2e})')
        ok &= same.item() < 1e-3 * (1 + abs(lp_1))
        # parameters identical on all ranks
        flat = torch.cat([p.detach().reshape(-1) for p in flow_dp.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        ok &= bool(torch.equal(flat, ref))
    if rank == 0:
        print('DP_FIT_OK' if ok else 'DP_FIT_MISMATCH')
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
