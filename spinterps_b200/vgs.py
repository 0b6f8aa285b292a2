"""Host-side variogram-string helpers (format of variograms/vgs.py:841-845:
``'%0.5f %s(%0.5f)'`` terms joined by ``' + '``)."""
from __future__ import annotations

import math

import numpy as np

from . import _lib


def check_full_nuggetness(in_model, min_vg_val):
    """Same decision as misc.py:1074-1105 of the reference: the sum of the
    texts before each 3-letter model name is the sill, the largest range is the
    range; nugget-only if either is <= min_vg_val."""
    in_model = str(in_model)
    if in_model == 'nan':
        return False
    sill_sum = 0.0
    rng_max = 0.0
    for submodel in in_model.split('+'):
        submodel = submodel.strip()
        sill_sum += float(submodel.split('(')[0].strip()[:-3].strip())
        rng_max = max(rng_max, float(submodel.split('(')[1].split(')')[0]))
    return bool((sill_sum <= min_vg_val) or (rng_max <= min_vg_val))


def get_vgs_cluster(vgs):
    """interp/vgclus.py:33-79: {vg string: step positions}, first-occurrence order."""
    clus = {}
    for i, vg in enumerate(vgs):
        clus.setdefault(vg, []).append(i)
    return {k: np.asarray(v, dtype=np.int64) for k, v in clus.items()}


def vg_abs_bound(vg_str, max_dist):
    """Upper bound of |vg(h)| and |sum(sill) - vg(h)| for 0 <= h <= max_dist.
    Used only to decide which systems need the explicit sum(lambda) check."""
    bound = 0.0
    for t, s, r in _lib.parse_vg_str(vg_str):
        name = _lib.VG_NAMES[t]
        if name == 'Rng':
            b = max_dist + abs(s)
        elif name == 'Pow':
            b = abs(s) * max(1.0, max_dist ** r) + abs(s)
        elif name == 'Hol':
            b = 1.25 * abs(s)
        else:
            b = abs(s)
        bound += b
    return bound if math.isfinite(bound) else float('inf')
