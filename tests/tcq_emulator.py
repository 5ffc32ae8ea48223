"""CPU emulation of csrc/b2f_flow_tcq.cu from the operand plan of torchflows_b200/_tcq.py (test infrastructure).

Executes exactly the dataflow the kernel implements -- materialise the source half, TF32 GEMM1, tanh, TF32 GEMM2 on the
folded weights, pre-affine, folded spline (the kernel's own header compiled for the host), post-affine, base-density
accumulation, final pass -- so that the operand folding can be checked against the oracle without a GPU."""
import math

import torch

from tests import hostmath as hm


def tf32_trunc(t):
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def tf32_round(t):
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def run(plan, x, logp_of_input=False, safe=False):
    """x: (B, D) CPU float32.  Returns y, log_det, log_prob as the kernel would write them."""
    B, D = x.shape
    Dh = D // 2
    half = (slice(0, Dh), slice(Dh, D))
    tile = x.clone()
    ld2 = torch.zeros(B)
    sq = torch.zeros(B)
    if logp_of_input:
        sq = ((plan.in_a.cpu() * x + plan.in_b.cpu()) ** 2).sum(-1)
    for layer in plan.layers:
        s, t = half[layer['src_half']], half[layer['tgt_half']]
        H = layer['H']
        if layer['src_pass']:
            tile[:, s] = layer['src_a'].cpu() * tile[:, s] + layer['src_b'].cpu()
        W1 = tf32_round(layer['W1'].cpu())
        pre = tf32_trunc(tile[:, s]).double() @ W1.double().T
        a2 = tf32_round(torch.tanh(pre.float() + layer['b1'].cpu()))
        M = layer['M'].cpu().double()                      # (Dh, 24, K2): tf32 weights, bias hi, bias lo
        g = torch.einsum('bh,eph->bep', a2.double(), M[:, :, :H]) + M[:, :, H] + M[:, :, H + 1]
        v = layer['pre_a'].cpu() * tile[:, t] + layer['pre_b'].cpu()
        out, l2 = hm.rqfast_g(v.reshape(-1), g.float().reshape(-1, 24), layer['boundary'], layer['inverse'], safe)
        out, l2 = out.reshape(B, Dh), l2.reshape(B, Dh)
        stored = layer['post_a'].cpu() * out + layer['post_b'].cpu()
        tile[:, t] = stored
        ld2 = ld2 + l2.sum(-1)
        if not logp_of_input:
            sq = sq + ((layer['fin_a'].cpu() * stored + layer['fin_b'].cpu()) ** 2).sum(-1)
    y = plan.fin_a.cpu() * tile + plan.fin_b.cpu()
    log_det = ld2 * math.log(2.0) + plan.const_ld.cpu()
    log_prob = -0.5 * sq + float(plan.const_lp) + log_det
    return y, log_det, log_prob
