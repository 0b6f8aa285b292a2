"""Build the REFERENCE's own implementation of the hot path into oracle/_ref/.

TEST INFRASTRUCTURE -- not part of the product.  Only tests/, __graft_entry__ and
bench.py's CPU arm (`--impl reference`, `cpu_baseline`) use what this produces.

    python oracle/build_ref.py [--force]

The reference (/root/reference, read-only) is Python + one Cython module and has no
setup.py.  The files on the gridded-interpolation path,

    cyth/interpmthds.pyx   (Cython, C++: the cpdef free functions and kriging classes)
    interp/steps.py        (SpInterpSteps._get_all_interp_outputs, the hot loop nest)
    interp/grps.py         (availability / neighbour groups)
    interp/vgclus.py       (variogram clusters)
    misc.py                (check_full_nuggetness, traceback_wrapper, ...)

are compiled FROM WHERE THEY LIE with Cython + g++ into extension modules under
oracle/_ref/spinterps/ (git-ignored; the .so files travel to the GPU box with the
gpurun snapshot, where /root/reference does not exist).  No reference source is copied
into the repository: the intermediate C/C++ files are written to a temporary directory
and removed.  oracle/ref_runner.py imports the result.
"""
import hashlib
import shutil
import subprocess
import sys
import sysconfig
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path('/root/reference')
OUT = HERE / '_ref'
PKG = OUT / 'spinterps'

# (source relative to the reference root, package directory, C++?)
UNITS = [
    ('cyth/interpmthds.pyx', 'cyth', True),
    ('interp/steps.py', 'interp', False),
    ('interp/grps.py', 'interp', False),
    ('interp/vgclus.py', 'interp', False),
    ('misc.py', '', False),
]


def _digest():
    h = hashlib.sha256()
    for rel, _, _ in UNITS:
        h.update(rel.encode())
        h.update((REF / rel).read_bytes())
    h.update(Path(__file__).read_bytes())
    return h.hexdigest()


def available():
    """The compiled reference is present (built here or shipped with the snapshot)."""
    suffix = sysconfig.get_config_var('EXT_SUFFIX')
    return all((PKG / sub / (Path(rel).stem + suffix)).exists() for rel, sub, _ in UNITS)


def build(force=False, verbose=False):
    if not REF.exists():
        return available()
    stamp = OUT / 'build.stamp'
    dig = _digest()
    if not force and available() and stamp.exists() and stamp.read_text() == dig:
        return True
    import numpy as np
    suffix = sysconfig.get_config_var('EXT_SUFFIX')
    inc = sysconfig.get_paths()['include']
    if PKG.exists():
        shutil.rmtree(PKG)
    tmp = Path(tempfile.mkdtemp(prefix='spx_refbuild_'))
    try:
        for rel, sub, cplus in UNITS:
            src = REF / rel
            stem = src.stem
            mod = 'spinterps.' + (sub + '.' if sub else '') + stem
            gen = tmp / (mod.replace('.', '_') + ('.cpp' if cplus else '.c'))
            cmd = [sys.executable, '-m', 'cython', '-3', '--module-name', mod, str(src), '-o',
                   str(gen)]
            if cplus:
                cmd.insert(4, '--cplus')
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError('cython failed on %s:\n%s%s' % (rel, r.stdout, r.stderr))
            dst = PKG / sub
            dst.mkdir(parents=True, exist_ok=True)
            so = dst / (stem + suffix)
            cc = ['g++' if cplus else 'gcc', '-shared', '-fPIC', '-O2', '-fwrapv',
                  '-fno-strict-aliasing', '-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION',
                  '-I', inc, '-I', np.get_include(), str(gen), '-o', str(so)]
            r = subprocess.run(cc, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError('compile failed on %s:\n%s' % (rel, r.stderr[-4000:]))
            if verbose:
                print('built', so.relative_to(HERE))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    stamp.write_text(dig)
    return True


if __name__ == '__main__':
    ok = build(force='--force' in sys.argv, verbose=True)
    print('oracle/_ref', 'ready' if ok else 'unavailable (no /root/reference, nothing prebuilt)')
