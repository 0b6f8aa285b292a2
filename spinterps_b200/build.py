"""Build libspx_b200.so in-tree with nvcc for sm_100a.

    python -m spinterps_b200.build [--force]

The shared library lands in spinterps_b200/lib/ (git-ignored, travels to the GPU
box with the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / 'csrc'
LIBDIR = PKG / 'lib'
LIB = LIBDIR / 'libspx_b200.so'
SOURCES = ['spx_basic.cu', 'spx_solve.cu', 'spx_gemm.cu', 'spx_misc.cu', 'spx_nrst.cu',
           'spx_plan.cu', 'spx_prep.cu', 'spx_chunk.cu', 'spx_pack.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=default',
    '-Xptxas', '-v', '--fmad=true']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _digest():
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob('*.cu')) + list(CSRC.glob('*.cuh')) + list(CSRC.glob('*.h')) +
                    [ROOT / 'include' / 'spx_b200.h', Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def _src_digest(src):
    """Digest of one translation unit: its source, every header, the flags."""
    h = hashlib.sha256()
    for p in ([CSRC / src] + sorted(CSRC.glob('*.cuh')) + sorted(CSRC.glob('*.h')) +
              [ROOT / 'include' / 'spx_b200.h']):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    obj = LIBDIR / (src[:-3] + '.o')
    cmd = [_nvcc(), *NVCC_FLAGS, '-I', str(ROOT / 'include'), '-I', str(CSRC),
           '-c', str(CSRC / src), '-o', str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout, r.stderr


def build(force=False, verbose=False):
    """Compile the translation units that changed (in parallel) and link.  build.stamp
    holds the digest of the whole source set the shipped .so was built from; the
    per-object digests (*.o.sha) and the ptxas register / spill report (ptxas.log) are
    untracked build artefacts."""
    from concurrent.futures import ThreadPoolExecutor
    LIBDIR.mkdir(exist_ok=True)
    stamp = LIBDIR / 'build.stamp'
    dig = _digest()
    if (not force) and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    todo = []
    for src in SOURCES:
        obj = LIBDIR / (src[:-3] + '.o')
        sha = LIBDIR / (src[:-3] + '.o.sha')
        if force or not obj.exists() or not sha.exists() or sha.read_text() != _src_digest(src):
            todo.append(src)
    log = []
    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4) or 1) as ex:
        for src, rc, out, err in ex.map(_compile, todo):
            log.append('== %s\n%s' % (src, err))
            if rc != 0:
                sys.stderr.write(out + err)
                raise RuntimeError(f'nvcc failed on {src}')
            (LIBDIR / (src[:-3] + '.o.sha')).write_text(_src_digest(src))
    objs = [str(LIBDIR / (src[:-3] + '.o')) for src in SOURCES]
    cmd = [_nvcc(), '-shared', '-o', str(LIB), *objs, '-lcudart']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('link failed')
    with open(LIBDIR / 'ptxas.log', 'a' if (LIBDIR / 'ptxas.log').exists() and not force else 'w') as f:
        f.write('\n'.join(log))
    stamp.write_text(dig)
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
