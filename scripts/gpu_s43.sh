#!/bin/bash
O=gpurun_out
CUDA_VISIBLE_DEVICES=0 python scripts/probe_e2e_delta.py > $O/s43_p0.log 2>&1 &
CUDA_VISIBLE_DEVICES=1 python scripts/probe_e2e_delta.py > $O/s43_p1.log 2>&1 &
wait
head -12 $O/s43_p0.log; echo ----; head -12 $O/s43_p1.log
