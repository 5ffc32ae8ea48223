"""Operand layout of the MADE spline tensor-core kernel (csrc/b2f_flow_tcm.cu, include/b2f.h B2F_FLAG_TCM_OPERANDS).

One-pass direction of MaskedAutoregressiveRQNSF-style programs (ElementwiseAffine / ActNorm, ReversePermutation,
MADE-conditioned RQS layers: architectures.py:147-163, layers_base.py:202-211, transforms.py:184-266 of the reference).
As in torchflows_b200/_tcq.py everything batch-independent is precomputed: the MADE masks are multiplied into both weight
matrices, the permutation becomes a column / element order of the weights, runs of elementwise layers become per-column
affine maps that ride on the write-back of the layer before them (or one materialisation pass in front of the first
layer), and the output layer is folded into the 24 columns per element of csrc/b2f_rqfast.cuh.
"""
import math
from typing import List, Optional, Sequence

import torch

from . import _native as N
from ._tcq import CPE, EPC, _elementwise_affine, canonical, fold_output_layer, round_tf32

MAGIC = 0x4D435442                              # 'BTCM'


def eligible(ops: Sequence, D: int) -> bool:
    """Mirror of try_launch_flow_tcm's conditions: spline MADE one-pass programs, D a multiple of 32 in [32, 128], hidden
    width <= 30, even number of flips."""
    if D % 32 != 0 or D < 32 or D > 128 or len(ops) > N.MAX_OPS:
        return False
    flip, n_m = False, 0
    for op in ops:
        if op.kind == N.OP_FLIP:
            flip = not flip
        elif op.kind == N.OP_ELEMENTWISE:
            if op.tkind not in (N.T_AFFINE_FWD, N.T_AFFINE_INV) or (op.flags & N.FLAG_ROW_BIAS):
                return False
        elif op.kind == N.OP_MADE:
            if op.tkind not in (N.T_RQ_FWD, N.T_RQ_INV) or op.n_bins != 8 or not (1 <= op.n_hidden <= 30):
                return False
            n_m += 1
        else:
            return False
    return (not flip) and 1 <= n_m <= 12


class Plan:
    def __init__(self):
        self.layer_blobs: List[torch.Tensor] = []
        self.program_blob: Optional[torch.Tensor] = None


def build_plan(ops: Sequence, D: int, base_loc: Optional[torch.Tensor], base_log_scale: Optional[torch.Tensor]) -> Plan:
    dev = None
    for op in ops:
        if op.leafs:
            dev = op.leafs[0].device
            break
    f32 = dict(device=dev, dtype=torch.float32)
    A, B = torch.ones(D, **f32), torch.zeros(D, **f32)          # pending affine map per PHYSICAL column
    pending = False
    const_ld = torch.zeros((), **f32)
    flip, last = False, None
    layers = []
    plan = Plan()
    with torch.no_grad():
        for op in ops:
            if op.kind == N.OP_FLIP:
                flip = not flip
            elif op.kind == N.OP_ELEMENTWISE:
                a, b, ld = _elementwise_affine(op, flip)
                const_ld = const_ld + ld
                if last is not None:            # rides on the write-back of the MADE layer before (it writes every column)
                    last['post_b'] = a * last['post_b'] + b
                    last['post_a'] = a * last['post_a']
                else:
                    B = a * B + b
                    A = a * A
                    pending = True
            else:
                W1, b1, W2, b2 = (x.detach().float() for x in op.leafs)
                m1, m2 = (c.to(device=dev, dtype=torch.float32) for c in op.consts[:2])
                W1, W2 = W1 * m1, W2 * m2                      # masks folded into the weight tiles (transforms.py:197-198)
                H = W1.shape[0]
                Wf, bf = fold_output_layer(W2, b2, D)         # (D, 24, H), (D, 24)
                if flip:                                       # physical column / element order is the reverse of the logical one
                    W1, Wf, bf = W1.flip(1), Wf.flip(0), bf.flip(0)
                layer = dict(src_pass=pending, H=H, K2=(H + 2 + 7) // 8 * 8, inverse=op.tkind == N.T_RQ_INV,
                             src_a=A.clone(), src_b=B.clone(), post_a=torch.ones(D, **f32), post_b=torch.zeros(D, **f32),
                             W1=W1, b1=b1, Wf=Wf, bf=bf)
                A, B, pending = torch.ones(D, **f32), torch.zeros(D, **f32), False
                last = layer
                layers.append(layer)
        ls = base_log_scale.detach().float() if base_log_scale is not None else torch.zeros(D, **f32)
        loc = base_loc.detach().float() if base_loc is not None else torch.zeros(D, **f32)
        inv_s = torch.exp(-ls)
        # after the last MADE layer nothing is pending (trailing elementwise layers rode on its write-back)
        fin_a, fin_b = A.clone(), B.clone()
        const_lp = -(0.5 * math.log(2 * math.pi) * D + ls.sum())
        for i, layer in enumerate(layers):
            if i == len(layers) - 1:
                layer['fin_a'] = inv_s
                layer['fin_b'] = -loc * inv_s
            else:
                layer['fin_a'] = torch.zeros(D, **f32)
                layer['fin_b'] = torch.zeros(D, **f32)
        for layer in layers:
            H, K2 = layer['H'], layer['K2']
            n_chunks = D // EPC
            hdr = torch.tensor([MAGIC, 0, int(layer['src_pass']), H, K2, n_chunks, D, 0], dtype=torch.int32,
                               device=dev).view(torch.float32)
            W1p = torch.zeros(32, D, **f32)
            W1p[:H] = round_tf32(layer['W1'])
            b1p = torch.zeros(32, **f32)
            b1p[:H] = layer['b1']
            M = torch.zeros(D, CPE, K2, **f32)
            M[:, :, :H] = round_tf32(layer['Wf'])
            b_hi = round_tf32(layer['bf'])
            M[:, :, H] = b_hi
            M[:, :, H + 1] = round_tf32(layer['bf'] - b_hi)
            w2c = M.reshape(n_chunks, (EPC * CPE) // 8, 8, K2 // 4, 4).permute(0, 1, 3, 2, 4).contiguous().reshape(-1)
            one, zero = torch.ones(D, **f32), torch.zeros(D, **f32)
            tp = torch.stack([one, zero, layer['post_a'], layer['post_b'], layer['fin_a'], layer['fin_b'], zero, zero],
                             dim=1).reshape(-1)
            sp = torch.stack([layer['src_a'], layer['src_b']], dim=1).reshape(-1)
            absum = layer['Wf'].abs().sum(dim=2) + layer['bf'].abs()
            misc = torch.stack([absum[:, 0:8].max(), absum[:, 8:16].max(), torch.zeros((), **f32), torch.zeros((), **f32)])
            plan.layer_blobs.append(torch.cat([hdr, canonical(W1p), b1p, w2c, tp, sp, misc]).contiguous())
        flags = torch.tensor([MAGIC, int(pending), 0, len(layers)], dtype=torch.int32, device=dev).view(torch.float32)
        consts = torch.stack([const_ld.reshape(()), const_lp.reshape(()).float(), torch.zeros((), **f32), torch.zeros((), **f32)])
        plan.program_blob = torch.cat([flags, consts, torch.stack([fin_a, fin_b], dim=1).reshape(-1),
                                       torch.stack([inv_s, -loc * inv_s], dim=1).reshape(-1)]).contiguous()
    return plan


def cached_plan(ops: Sequence, D: int, base_loc, base_log_scale) -> Plan:
    """Plan of a program, rebuilt when any parameter (or the base distribution) changed."""
    tensors = [t for op in ops for t in op.leafs] + [t for t in (base_loc, base_log_scale) if t is not None]
    ver = tuple((t.data_ptr(), t._version) for t in tensors) + tuple((op.kind, op.tkind) for op in ops)
    owner = next((op.owner for op in ops if op.kind == N.OP_MADE and op.owner is not None), None)
    cache = getattr(owner, '_b2f_cache', None) if owner is not None else None
    if cache is None:
        cache = {}
        if owner is not None:
            object.__setattr__(owner, '_b2f_cache', cache)
    key = ('tcm', base_loc is None, base_log_scale is None) + tuple(op.tkind for op in ops)
    hit = cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    plan = build_plan(ops, D, base_loc, base_log_scale)
    cache[key] = (ver, plan)
    return plan
