#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "idw or golden or 1000 or 2000 or seeded" > $O/s50_tests.log 2>&1; tail -3 $O/s50_tests.log
for d in 0 1; do SPX_GEMM_DIRECT=$d timeout 300 python scripts/probe_gemm_cfg.py c4 2>&1 | tail -1; done
for d in 0 2; do SPX_GEMM_DIRECT=$d timeout 300 python scripts/probe_gemm_cfg.py dense 2>&1 | tail -1; done
