"""spinterps_b200 -- B200-native (sm_100a) gridded interpolation hot path of
faizan90/spinterps: OK / SK / EDK kriging, IDW and nearest neighbour of station
time series onto a raster grid.

Public surface (mirrors the reference's):
    SpInterpMain         configuration + verify() + interpolate()
    SpInterpSteps        the per-chunk operator seam
    ChunkEngine          array-level engine (submit_chunk / interp_chunk)
    cyth                 drop-in for the reference's Cython free functions

Importing the package does not need a GPU; every compute call does (there is no
CPU fallback).
"""
from . import _lib  # noqa: F401

__all__ = ['SpInterpMain', 'SpInterpSteps', 'ChunkEngine', 'cyth']


def __getattr__(name):
    # lazy: torch / pandas are only imported when the classes are used
    if name == 'SpInterpMain':
        from .main import SpInterpMain
        return SpInterpMain
    if name == 'SpInterpSteps':
        from .steps import SpInterpSteps
        return SpInterpSteps
    if name == 'ChunkEngine':
        from .engine import ChunkEngine
        return ChunkEngine
    if name == 'cyth':
        import importlib
        return importlib.import_module('.cyth', __name__)
    raise AttributeError(name)
