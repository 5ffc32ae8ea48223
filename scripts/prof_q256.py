"""Small driver for ncu / timing experiments on the Q256 workload: python scripts/prof_q256.py [rows] [reps]."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import build_flow  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    preset = os.environ.get('PROF_PRESET', 'CouplingRQNSF')
    D = int(os.environ.get('PROF_D', '256'))
    dev = torch.device('cuda:0')
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(rows, D, device=dev, generator=g)
    z = torch.randn(rows, D, device=dev, generator=g)
    flow = build_flow(preset, D, dev, init_rows=x[:65536])
    with torch.no_grad():
        for _ in range(2):
            flow.log_prob(x)
            flow._sample_from_base(z, no_grad=True)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        for _ in range(reps):
            flow.log_prob(x)
        e[1].record()
        for _ in range(reps):
            flow._sample_from_base(z, no_grad=True)
        e[2].record()
        torch.cuda.synchronize()
    print(f'{preset}({D}) rows={rows}: log_prob {e[0].elapsed_time(e[1]) / reps:.3f} ms, sample {e[1].elapsed_time(e[2]) / reps:.3f} ms',
          flush=True)


if __name__ == '__main__':
    main()
