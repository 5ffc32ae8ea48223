"""ctypes front end of oracle/b2f_oracle.c (plain-C CPU oracle).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libb2f_oracle.so')
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            import subprocess
            os.makedirs(os.path.dirname(_SO), exist_ok=True)
            subprocess.run(['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared',
                            os.path.join(_HERE, 'b2f_oracle.c'), '-o', _SO, '-lm'], check=True)
        _lib = ctypes.CDLL(_SO)
        _lib.b2f_oracle_exp.restype = ctypes.c_float
        _lib.b2f_oracle_exp.argtypes = [ctypes.c_float]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def rq(x: torch.Tensor, h: torch.Tensor, n_bins: int, boundary: float, inverse: bool):
    """x: (*batch, *event), h: (*batch, *event, 3*n_bins-1).  Returns out, per-element log-det, k (int32)."""
    xs = np.ascontiguousarray(x.detach().cpu().numpy(), dtype=np.float32)
    hs = np.ascontiguousarray(h.detach().cpu().numpy(), dtype=np.float32)
    out = np.empty_like(xs)
    ld = np.empty_like(xs)
    k = np.empty(xs.shape, dtype=np.int32)
    lib().b2f_oracle_rq(_p(xs), _p(hs), _p(out), _p(ld), _p(k), ctypes.c_int64(xs.size), ctypes.c_int(n_bins),
                        ctypes.c_float(boundary), ctypes.c_int(int(inverse)))
    return torch.from_numpy(out), torch.from_numpy(ld), torch.from_numpy(k)


def rq_knots(u: torch.Tensor, boundary: float):
    us = np.ascontiguousarray(u.detach().cpu().numpy(), dtype=np.float32)
    nb = us.shape[-1]
    kn = np.empty(us.shape[:-1] + (nb + 1,), dtype=np.float32)
    lib().b2f_oracle_rq_knots(_p(us), _p(kn), ctypes.c_int64(us.size // nb), ctypes.c_int(nb), ctypes.c_float(boundary))
    return torch.from_numpy(kn)


def affine(x: torch.Tensor, h: torch.Tensor, inverse: bool):
    xs = np.ascontiguousarray(x.detach().cpu().numpy(), dtype=np.float32)
    hs = np.ascontiguousarray(h.detach().cpu().numpy(), dtype=np.float32)
    out = np.empty_like(xs)
    ld = np.empty_like(xs)
    lib().b2f_oracle_affine(_p(xs), _p(hs), _p(out), _p(ld), ctypes.c_int64(xs.size), ctypes.c_int(int(inverse)))
    return torch.from_numpy(out), torch.from_numpy(ld)


def exp_det(t: float) -> float:
    return float(lib().b2f_oracle_exp(ctypes.c_float(t)))
