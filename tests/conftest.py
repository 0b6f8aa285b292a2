import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

os.environ.setdefault('MASTER_ADDR', '127.0.0.1')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    """Build libspx_b200.so if it is missing or stale (nvcc cross-compiles
    without a GPU; ~1 min the first time, cached by a source digest)."""
    from spinterps_b200 import build
    build.build()


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a host without a CUDA device skips the gpu-marked tests instead of
    failing them (the product itself still refuses to run there: no CPU fallback)."""
    try:
        from spinterps_b200 import _lib, build
        build.build()                      # no-op when the shipped library is current
        have_gpu = _lib.load().spx_device_count() >= 1
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device visible')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Parity report of the GPU tests (floored and un-floored relative errors per case)."""
    try:
        from tests import test_gpu_engine as tge
        if tge.PARITY_REPORT:
            import json
            out = ROOT / 'gpurun_out'
            out.mkdir(exist_ok=True)
            (out / 'parity_report.json').write_text(json.dumps(tge.PARITY_REPORT, indent=1))
    except Exception:
        pass
