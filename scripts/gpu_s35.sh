#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_main.py -m gpu -x -q > $O/s35_tests.log 2>&1; tail -25 $O/s35_tests.log
timeout 300 python scripts/probe_e2e_delta.py > $O/s35_probe.log 2>&1; head -12 $O/s35_probe.log
