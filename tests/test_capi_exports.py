"""The C-ABI library loads and exports every symbol include/spx_b200.h declares;
host-only entry points work; compute entry points fail loudly without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from spinterps_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def _declared_functions():
    src = (ROOT / 'include' / 'spx_b200.h').read_text()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(spx_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported():
    names = _declared_functions()
    assert len(names) >= 25
    lib = C.CDLL(str(_lib.LIB_PATH))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding knows every one of them
    assert sorted(_lib.EXPORTED) == names


def test_version_and_error_channel():
    lib = _lib.load()
    assert lib.spx_version() >= 100
    assert isinstance(lib.spx_last_error(), bytes)
    assert lib.spx_device_count() >= 0


def test_parse_vg_str_matches_reference_grammar():
    # cyth/interpmthds.pyx:174-184
    assert _lib.parse_vg_str('0.1 Nug(0.0) + 0.9 Sph(20000)') == [(1, 0.1, 1e-5), (2, 0.9, 20000.0)]
    assert _lib.parse_vg_str('0.1 Nug(0.0) + 0.9 Sph(20000)', clamp_range=False)[0][2] == 0.0
    assert _lib.parse_vg_str(' 1.5 Exp(3e4)+2 Gau(10) ') == [(3, 1.5, 3e4), (5, 2.0, 10.0)]
    names = ['Rng', 'Nug', 'Sph', 'Exp', 'Lin', 'Gau', 'Pow', 'Hol']
    for i, n in enumerate(names):
        assert _lib.parse_vg_str(f'1.0 {n}(5.0)')[0][0] == i
    for bad in ['nan', '0.1  Nug(0.0)', '0.1 Foo(1)', '0.1 Nug', 'x Nug(1)', '0.1 Nug(y)', '']:
        with pytest.raises(_lib.SpxError):
            _lib.parse_vg_str(bad)
    with pytest.raises(_lib.SpxError):   # more than SPX_VG_MAX_TERMS nested terms
        _lib.parse_vg_str(' + '.join(['0.1 Sph(10)'] * 9))


def test_coef_offset_is_a_bijection_on_a_tile():
    lib = _lib.load()
    kpad = 16
    offs = {lib.spx_coef_offset(r, c, kpad) for r in range(512) for c in range(kpad)}
    assert offs == set(range(512 * kpad))
    # 8x4 fragment blocks are contiguous and lane-major
    assert [lib.spx_coef_offset(r, c, kpad) for r in range(2) for c in range(4)] == list(range(8))
    assert lib.spx_coef_offset(8, 0, kpad) == 32
    assert lib.spx_coef_offset(0, 4, kpad) == 256 * 4
    assert lib.spx_coef_offset(256, 0, kpad) == 256 * kpad


def test_struct_layouts_match_header_sizes(tmp_path):
    """sizeof / last-field offset of every struct of the header, as the C compiler
    sees them, equal the ctypes mirrors (a drifted mirror would corrupt arguments)."""
    import subprocess
    structs = {'spx_vg': 'ranges', 'spx_systems': 'max_m', 'spx_rhs': 'coef_row_major',
               'spx_downdate': 'base_f', 'spx_dd_plan': 'n_bytes', 'spx_gemm': 'quad_slot',
               'spx_multivg': 'all_fast', 'spx_local': 'slot', 'spx_nrst': 'u_end',
               'spx_fast_cfg': 'sparse', 'spx_pack_row': 'n_nan', 'spx_fast_result': 'host_ms',
               'spx_sparse_cov': 'blk'}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "spx_b200.h"', 'int main(void){']
    for name, last in structs.items():
        src.append(f'printf("{name} %zu %zu\\n", sizeof({name}), offsetof({name}, {last}));')
    src.append('return 0;}')
    c = tmp_path / 'sz.c'
    c.write_text('\n'.join(src))
    exe = tmp_path / 'sz'
    subprocess.run(['gcc', '-I', str(ROOT / 'include'), str(c), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        name, size, off = line.split()
        ct = getattr(_lib, name)
        assert C.sizeof(ct) == int(size), (name, C.sizeof(ct), size)
        assert getattr(ct, structs[name]).offset == int(off), (name, 'last field offset')
    assert _lib.VG_DTYPE.itemsize == C.sizeof(_lib.spx_vg)


def test_compute_fails_loudly_without_gpu():
    if _lib.load().spx_device_count() > 0:
        pytest.skip('a GPU is visible')
    from spinterps_b200 import cyth
    x = np.array([0.0, 3.0])
    y = np.array([0.0, 4.0])
    with pytest.raises(_lib.SpxError):
        cyth.fill_dists_2d_mat(x, y, x, y, np.full((2, 2), np.nan))
    with pytest.raises(_lib.SpxError):
        _lib.require_gpu()
    from spinterps_b200.engine import ChunkEngine
    with pytest.raises(_lib.SpxError):
        ChunkEngine()


def test_product_does_not_import_the_oracle():
    for path in (ROOT / 'spinterps_b200').rglob('*.py'):
        assert 'oracle' not in path.read_text().replace('see oracle/spinterp_oracle.py', ''), path
