"""Build recipes for the native pieces of this repo (all outputs are git-ignored *.so files that
travel to the GPU box with the snapshot).

  product : torchflows_b200/lib/libb2f.so        nvcc, sm_100a, the C-ABI library (include/b2f.h)
  oracle  : oracle/_build/libb2f_oracle.so       gcc, plain-C CPU oracle (test infrastructure)
  hostmath: tests/_build/libb2f_hostmath.so      g++, host build of csrc/b2f_math.cuh (test infrastructure)
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')

PRODUCT_SO = os.path.join(ROOT, 'torchflows_b200', 'lib', 'libb2f.so')
ORACLE_SO = os.path.join(ROOT, 'oracle', '_build', 'libb2f_oracle.so')
HOSTMATH_SO = os.path.join(ROOT, 'tests', '_build', 'libb2f_hostmath.so')


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print('+', ' '.join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build_product(force=False, verbose_ptxas=False):
    csrc = os.path.join(ROOT, 'torchflows_b200', 'csrc')
    cus = sorted(glob.glob(os.path.join(csrc, '*.cu')))
    deps = cus + glob.glob(os.path.join(csrc, '*.cuh')) + glob.glob(os.path.join(ROOT, 'include', '*.h'))
    if not force and not _newer(PRODUCT_SO, deps):
        return PRODUCT_SO
    os.makedirs(os.path.dirname(PRODUCT_SO), exist_ok=True)
    objs = []
    procs = []
    headers = [d for d in deps if not d.endswith('.cu')]
    for cu in cus:
        obj = os.path.join(os.path.dirname(PRODUCT_SO), os.path.basename(cu)[:-3] + '.o')
        objs.append(obj)
        if not force and not verbose_ptxas and not _newer(obj, [cu] + headers):
            continue                      # this translation unit is up to date
        cmd = [NVCC, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
               '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include'), '-c', cu, '-o', obj]
        if verbose_ptxas:
            cmd += ['-Xptxas', '-v']
        print('+', ' '.join(cmd), flush=True)
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError('nvcc failed')
    _run([NVCC, '-shared', '-o', PRODUCT_SO] + objs + ['-cudart', 'static'])
    return PRODUCT_SO


def build_oracle(force=False):
    src = os.path.join(ROOT, 'oracle', 'b2f_oracle.c')
    if force or _newer(ORACLE_SO, [src]):
        os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
        _run(['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared', src, '-o', ORACLE_SO, '-lm'])
    return ORACLE_SO


def build_hostmath(force=False):
    src = os.path.join(ROOT, 'tests', 'host_math.cpp')
    hdrs = [os.path.join(ROOT, 'torchflows_b200', 'csrc', n) for n in ('b2f_math.cuh', 'b2f_rqfast.cuh')]
    if force or _newer(HOSTMATH_SO, [src] + hdrs):
        os.makedirs(os.path.dirname(HOSTMATH_SO), exist_ok=True)
        _run(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-x', 'c++', src, '-o', HOSTMATH_SO])
    return HOSTMATH_SO


if __name__ == '__main__':
    what = sys.argv[1:] or ['product', 'oracle', 'hostmath']
    if 'oracle' in what:
        build_oracle(force=True)
    if 'hostmath' in what:
        build_hostmath(force=True)
    if 'product' in what:
        build_product(force=True, verbose_ptxas='-v' in what)
