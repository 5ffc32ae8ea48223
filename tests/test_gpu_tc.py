"""Tensor-core building blocks in isolation (tcgen05 / TMEM / UMMA descriptors, csrc/b2f_umma.cuh)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tf32(t):
    """tf32 keeps 10 mantissa bits; the tensor core ignores the low 13 bits of the fp32 operand."""
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize('N,K', [(16, 8), (32, 128), (192, 24), (240, 24), (256, 32), (64, 64)])
def test_umma_gemm_tile(N, K):
    from torchflows_b200 import _native as N_
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)
    C = N_.debug_umma_gemm(A, B)
    torch.cuda.synchronize()
    ref = tf32(A).double() @ tf32(B).double().T
    err = (C.double() - ref).abs().max().item()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    if err > 1e-3 * ref.abs().max().item():
        # leave a trace for offline layout forensics: one-hot probes
        probes = {}
        for (m, k) in ((0, 0), (1, 0), (8, 0), (0, 1), (0, 4), (0, 8), (9, 5)):
            Ah = torch.zeros(128, K, device=dev)
            Ah[m, k] = 1.0
            Bh = (torch.arange(N * K, device=dev, dtype=torch.float32).reshape(N, K) + 1)
            probes[(m, k)] = N_.debug_umma_gemm(Ah, Bh).cpu()
        torch.save({'C': C.cpu(), 'ref': ref.cpu(), 'probes': probes}, os.path.join(ROOT, 'gpurun_out', f'umma_debug_{N}_{K}.pt'))
    assert err <= 1e-3 * ref.abs().max().item(), err
