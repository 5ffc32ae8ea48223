"""Operand layout of the affine-coupling tensor-core kernel (csrc/b2f_flow_tca.cu, include/b2f.h B2F_FLAG_TCA_OPERANDS).

The kernel runs a whole RealNVP / NICE-style program (ElementwiseAffine / ActNorm, ReversePermutation, affine or shift
coupling layers: architectures.py:57-96 of the reference) with several 128-row tiles in flight per SM.  Like
torchflows_b200/_tcq.py it wants everything batch-independent precomputed: runs of elementwise layers folded into
per-column affine maps applied where a column is touched anyway, the permutation folded into the weight order, and both
conditioner layers as UMMA operands -- here split into tf32 hi / lo parts, because affine flows need fp32-faithful
conditioners: every product is evaluated as hi*hi + lo*hi + hi*lo on the tensor cores.
"""
import os
from typing import List, Optional, Sequence

import torch

from . import _native as N
from ._tcq import _elementwise_affine, canonical, round_tf32

MAGIC = 0x41435442                              # 'BTCA'
HDR = 8


def eligible(ops: Sequence, D: int) -> bool:
    """Mirror of try_launch_flow_tca's conditions: affine / shift coupling programs, D a multiple of 32 in [32, 128],
    hidden width <= 31, both halves transformed, even number of flips."""
    if D % 32 != 0 or D < 32 or D > 128 or len(ops) > N.MAX_OPS:
        return False
    if D > 64 and not os.environ.get('B2F_TCA_ANY'):      # fewer than four tile pipelines fit: the rows kernel is faster
        return False
    flip, written, n_c = False, set(), 0
    for op in ops:
        if op.kind == N.OP_FLIP:
            flip = not flip
        elif op.kind == N.OP_ELEMENTWISE:
            if op.tkind not in (N.T_AFFINE_FWD, N.T_AFFINE_INV) or (op.flags & N.FLAG_ROW_BIAS):
                return False
        elif op.kind == N.OP_COUPLING:
            if op.flags & N.FLAG_ROW_BIAS:              # context-conditioned layer: generic kernel
                return False
            if op.tkind not in (N.T_AFFINE_FWD, N.T_AFFINE_INV, N.T_SHIFT_ADD, N.T_SHIFT_SUB) or not (1 <= op.n_hidden <= 31):
                return False
            written.add(0 if flip else 1)
            n_c += 1
        else:
            return False
    return (not flip) and 1 <= n_c <= 8 and written == {0, 1}


def _split(t: torch.Tensor):
    hi = round_tf32(t)
    return hi, round_tf32(t - hi)


class Plan:
    def __init__(self):
        self.layer_blobs: List[torch.Tensor] = []
        self.program_blob: Optional[torch.Tensor] = None


def build_plan(ops: Sequence, D: int, base_loc: Optional[torch.Tensor], base_log_scale: Optional[torch.Tensor]) -> Plan:
    import math
    dev = None
    for op in ops:
        if op.leafs:
            dev = op.leafs[0].device
            break
    Dh = D // 2
    f32 = dict(device=dev, dtype=torch.float32)
    A, B = torch.ones(D, **f32), torch.zeros(D, **f32)
    pending = [False, False]
    const_ld = torch.zeros((), **f32)
    flip, last = False, None
    layers = []
    half = (slice(0, Dh), slice(Dh, D))
    plan = Plan()
    with torch.no_grad():
        for op in ops:
            if op.kind == N.OP_FLIP:
                flip = not flip
            elif op.kind == N.OP_ELEMENTWISE:
                a, b, ld = _elementwise_affine(op, flip)
                const_ld = const_ld + ld
                for h in (0, 1):
                    c = half[h]
                    if last is not None and last['tgt_half'] == h:      # rides on the write-back of the layer before
                        last['post_b'] = a[c] * last['post_b'] + b[c]
                        last['post_a'] = a[c] * last['post_a']
                    else:
                        B[c] = a[c] * B[c] + b[c]
                        A[c] = a[c] * A[c]
                        pending[h] = True
            else:
                s = 1 if flip else 0
                t = 1 - s
                W1, b1, W2, b2 = (x.detach().float() for x in op.leafs)
                H = W1.shape[0]
                P = 2 if op.tkind in (N.T_AFFINE_FWD, N.T_AFFINE_INV) else 1
                w2 = W2.reshape(Dh, P, H)
                bb = b2.reshape(Dh, P)
                if flip:                      # physical order of the source / target columns is the reverse of the logical one
                    W1, w2, bb = W1.flip(1), w2.flip(0), bb.flip(0)
                layer = dict(src_half=s, tgt_half=t, src_pass=pending[s], has_pre=pending[t], H=H, P=P, tkind=op.tkind,
                             src_a=A[half[s]].clone(), src_b=B[half[s]].clone(),
                             pre_a=A[half[t]].clone(), pre_b=B[half[t]].clone(),
                             post_a=torch.ones(Dh, **f32), post_b=torch.zeros(Dh, **f32), W1=W1, b1=b1, w2=w2, bb=bb)
                A[:] = 1.0
                B[:] = 0.0
                pending = [False, False]
                last = layer
                layers.append(layer)
        ls = base_log_scale.detach().float() if base_log_scale is not None else torch.zeros(D, **f32)
        loc = base_loc.detach().float() if base_loc is not None else torch.zeros(D, **f32)
        inv_s = torch.exp(-ls)
        fin_a, fin_b = A.clone(), B.clone()
        final_pass = list(pending)
        const_lp = -(0.5 * math.log(2 * math.pi) * D + ls.sum())
        last_writer = {}
        for i, layer in enumerate(layers):
            last_writer[layer['tgt_half']] = i
        for i, layer in enumerate(layers):
            c = half[layer['tgt_half']]
            if last_writer[layer['tgt_half']] == i:
                layer['fin_a'] = A[c] * inv_s[c]
                layer['fin_b'] = (B[c] - loc[c]) * inv_s[c]
            else:
                layer['fin_a'] = torch.zeros(Dh, **f32)
                layer['fin_b'] = torch.zeros(Dh, **f32)
        for i, layer in enumerate(layers):
            H, P = layer['H'], layer['P']
            N1 = (H + 15) // 16 * 16
            K2 = (H + 1 + 7) // 8 * 8
            N2 = (Dh * P + 15) // 16 * 16
            inverse = layer['tkind'] in (N.T_AFFINE_INV, N.T_SHIFT_SUB)
            w2, bb = layer['w2'].clone(), layer['bb'].clone()
            post_a, post_b = layer['post_a'], layer['post_b']
            folded = 0
            if P == 2:
                # the elementwise layers that follow ride on the conditioner's own output layer: with alpha = exp(c0 + u0 / 2),
                #   forward:  post_a (alpha v + u1) + post_b = (post_a alpha) v + (post_a u1 + post_b)
                #   inverse:  post_a (v - u1) / alpha + post_b = (v - u1) / (alpha / post_a) + post_b
                # i.e. u0 is shifted by +-2 log post_a (post_a > 0: a product of exp(.) scales) and, forward, u1 is mapped
                # affinely -- both linear in the output layer, so they go into its weights; the layer's log-det is then
                # +-(c0 + u0' / 2) -+ log post_a, the second term a constant of the program.
                folded = 1
                lpa = torch.log(post_a)
                const_ld = const_ld - lpa.sum()
                if inverse:
                    bb[:, 0] = bb[:, 0] - 2.0 * lpa
                else:
                    bb[:, 0] = bb[:, 0] + 2.0 * lpa
                    w2[:, 1, :] = w2[:, 1, :] * post_a[:, None]
                    bb[:, 1] = bb[:, 1] * post_a + post_b
            fin_mode = 2 if last_writer[layer['tgt_half']] == i else 0
            bits = P | (int(inverse) << 8) | (int(layer['has_pre']) << 9) | (fin_mode << 10) | (folded << 12)
            hdr = torch.tensor([MAGIC, layer['src_half'], int(layer['src_pass']), H, N1, K2, Dh, bits],
                               dtype=torch.int32, device=dev).view(torch.float32)
            W1p = torch.zeros(N1, Dh, **f32)
            W1p[:H] = layer['W1']
            w1h, w1l = _split(W1p)
            b1p = torch.zeros(32, **f32)
            b1p[:H] = layer['b1']
            M = torch.zeros(N2, K2, **f32)                     # row = element * P + parameter; column H = bias (times the 1 in A2)
            M[:Dh * P, :H] = w2.reshape(Dh * P, H)
            M[:Dh * P, H] = bb.reshape(Dh * P)
            w2h, w2l = _split(M)
            tp = torch.stack([layer['pre_a'], layer['pre_b'], post_a, post_b, layer['fin_a'],
                              layer['fin_b'], torch.zeros(Dh, **f32), torch.zeros(Dh, **f32)], dim=1).reshape(-1)
            sp = torch.stack([layer['src_a'], layer['src_b']], dim=1).reshape(-1)
            blob = torch.cat([hdr, canonical(w1h), canonical(w1l), b1p, canonical(w2h), canonical(w2l), tp, sp]).contiguous()
            plan.layer_blobs.append(blob)
        flags = torch.tensor([MAGIC, int(final_pass[0]), int(final_pass[1]), len(layers)], dtype=torch.int32,
                             device=dev).view(torch.float32)
        consts = torch.stack([const_ld.reshape(()), const_lp.reshape(()).float(), torch.zeros((), **f32), torch.zeros((), **f32)])
        plan.program_blob = torch.cat([flags, consts, torch.stack([fin_a, fin_b], dim=1).reshape(-1),
                                       torch.stack([inv_s, -loc * inv_s], dim=1).reshape(-1)]).contiguous()
    return plan


def cached_plan(ops: Sequence, D: int, base_loc, base_log_scale) -> Plan:
    """Plan of a program, rebuilt when any parameter (or the base distribution) changed."""
    tensors = [t for op in ops for t in op.leafs] + [t for t in (base_loc, base_log_scale) if t is not None]
    ver = tuple((t.data_ptr(), t._version) for t in tensors) + tuple((op.kind, op.tkind) for op in ops)
    owner = next((op.owner for op in ops if op.kind == N.OP_COUPLING and op.owner is not None), None)
    cache = getattr(owner, '_b2f_cache', None) if owner is not None else None
    if cache is None:
        cache = {}
        if owner is not None:
            object.__setattr__(owner, '_b2f_cache', cache)
    key = ('tca', base_loc is None, base_log_scale is None) + tuple(op.tkind for op in ops)
    hit = cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    plan = build_plan(ops, D, base_loc, base_log_scale)
    cache[key] = (ver, plan)
    return plan
