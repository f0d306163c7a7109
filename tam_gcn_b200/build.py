"""In-tree build of libtamgcn.so (hand-written sm_100a CUDA behind the C-ABI of include/tamgcn.h).

    python -m tam_gcn_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  The .so is written next to the sources
(tam_gcn_b200/lib/libtamgcn.so): git-ignored, but it travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, 'csrc')
OUT_DIR = os.path.join(PKG, 'lib')
OBJ_DIR = os.path.join(OUT_DIR, 'obj')
LIB = os.path.join(OUT_DIR, 'libtamgcn.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-I', os.path.join(ROOT, 'include')]


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: libtamgcn.so cannot be built')
    return exe


def sources():
    return sorted(os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith('.cu'))


def _deps_mtime():
    hs = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith(('.cuh', '.h'))]
    hs.append(os.path.join(ROOT, 'include', 'tamgcn.h'))
    return max(os.path.getmtime(h) for h in hs)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link libtamgcn.so.  Returns the library path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_m = _deps_mtime()
    jobs = []
    objs = []
    for s in sources():
        o = os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + '.o')
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError('nvcc failed on %s' % s)
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static',
                                                      '-Xlinker', '--no-undefined']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link of libtamgcn.so failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
