#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "idw or golden or 1000 or 2000 or seeded or est_vars" > $O/s51_tests.log 2>&1; tail -3 $O/s51_tests.log
for d in 0 4; do SPX_GEMM_KSPLIT=$d timeout 300 python scripts/probe_gemm_cfg.py c4 2>&1 | tail -1; done
