"""ctypes front end of tests/host_math.cpp: the kernels' per-element math compiled for the host (test only)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libb2f_hostmath.so')
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, 'host_math.cpp')
        hdrs = [os.path.join(_HERE, '..', 'torchflows_b200', 'csrc', n) for n in ('b2f_math.cuh', 'b2f_rqfast.cuh')]
        if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in [src] + hdrs):
            os.makedirs(os.path.dirname(_SO), exist_ok=True)
            subprocess.run(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-x', 'c++', src,
                            '-o', _SO], check=True)
        _lib = ctypes.CDLL(_SO)
        _lib.hm_exp_det.restype = ctypes.c_float
        _lib.hm_exp_det.argtypes = [ctypes.c_float]
    return _lib


def _np(t):
    return np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def rq(x, h, n_bins, boundary, inverse, templated=True):
    xs, hs = _np(x), _np(h)
    out, ld, k = np.empty_like(xs), np.empty_like(xs), np.empty(xs.shape, dtype=np.int32)
    lib().hm_rq(_p(xs), _p(hs), _p(out), _p(ld), _p(k), ctypes.c_int64(xs.size), ctypes.c_int(n_bins),
                ctypes.c_float(boundary), ctypes.c_int(int(inverse)), ctypes.c_int(int(templated)))
    return torch.from_numpy(out), torch.from_numpy(ld), torch.from_numpy(k)


def rqfast(x, h, boundary, inverse, safe=False, ny=0):
    """Folded-parameter spline of the tensor-core kernel (csrc/b2f_rqfast.cuh); h: (..., 23) raw parameters."""
    xs, hs = _np(x), _np(h)
    out, ld = np.empty_like(xs), np.empty_like(xs)
    lib().hm_rqfast(_p(xs), _p(hs), _p(out), _p(ld), ctypes.c_int64(xs.size), ctypes.c_float(boundary),
                    ctypes.c_int(int(inverse)), ctypes.c_int(int(safe)), ctypes.c_int(ny))
    return torch.from_numpy(out), torch.from_numpy(ld)


def rqfast_g(x, g, boundary, inverse, safe=False):
    """Same, the 24 folded columns per element given directly; returns out and the log-det in log2 units."""
    xs, gs = _np(x), _np(g)
    out, ld2 = np.empty_like(xs), np.empty_like(xs)
    lib().hm_rqfast_g(_p(xs), _p(gs), _p(out), _p(ld2), ctypes.c_int64(xs.size), ctypes.c_float(boundary),
                      ctypes.c_int(int(inverse)), ctypes.c_int(int(safe)))
    return torch.from_numpy(out), torch.from_numpy(ld2)


def rq_backward(x, h, gz, gl, n_bins, boundary):
    xs, hs, gzs, gls = _np(x), _np(h), _np(gz), _np(gl)
    dv, dh = np.empty_like(xs), np.empty_like(hs)
    lib().hm_rq_backward(_p(xs), _p(hs), _p(gzs), _p(gls), _p(dv), _p(dh), ctypes.c_int64(xs.size),
                         ctypes.c_int(n_bins), ctypes.c_float(boundary))
    return torch.from_numpy(dv), torch.from_numpy(dh)


def rq_backward_fast(x, h, gz, gl, boundary):
    """rqf::backward_fwd (csrc/b2f_rqfast.cuh): the wide-conditioner kernel's backward epilogue, n_bins = 8."""
    xs, hs, gzs, gls = _np(x), _np(h), _np(gz), _np(gl)
    dv, dh = np.empty_like(xs), np.empty_like(hs)
    lib().hm_rq_backward_fast(_p(xs), _p(hs), _p(gzs), _p(gls), _p(dv), _p(dh), ctypes.c_int64(xs.size),
                              ctypes.c_float(boundary))
    return torch.from_numpy(dv), torch.from_numpy(dh)


def rq_backward_inv(z, h, gx, gl, n_bins, boundary):
    zs, hs, gxs, gls = _np(z), _np(h), _np(gx), _np(gl)
    dz, dh = np.empty_like(zs), np.empty_like(hs)
    lib().hm_rq_backward_inv(_p(zs), _p(hs), _p(gxs), _p(gls), _p(dz), _p(dh), ctypes.c_int64(zs.size),
                             ctypes.c_int(n_bins), ctypes.c_float(boundary))
    return torch.from_numpy(dz), torch.from_numpy(dh)


def affine(x, h, inverse):
    xs, hs = _np(x), _np(h)
    out, ld = np.empty_like(xs), np.empty_like(xs)
    lib().hm_affine(_p(xs), _p(hs), _p(out), _p(ld), ctypes.c_int64(xs.size), ctypes.c_int(int(inverse)))
    return torch.from_numpy(out), torch.from_numpy(ld)


def affine_backward(x, h, gz, gl, inverse):
    xs, hs, gzs, gls = _np(x), _np(h), _np(gz), _np(gl)
    dx, dh = np.empty_like(xs), np.empty_like(hs)
    lib().hm_affine_backward(_p(xs), _p(hs), _p(gzs), _p(gls), _p(dx), _p(dh), ctypes.c_int64(xs.size),
                             ctypes.c_int(int(inverse)))
    return torch.from_numpy(dx), torch.from_numpy(dh)
