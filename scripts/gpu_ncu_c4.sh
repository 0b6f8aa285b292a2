#!/bin/bash
# ncu --set full of one contraction launch of config 4 (IDW, kpad 2008): one pass over the full
# K against one of the four K passes
O=gpurun_out
SPX_GEMM_KSPLIT=0 ncu --set full --clock-control none --import-source on -k regex:k_estimate_gemm -s 1 -c 1 -o $O/r2c_gemm_c4_onepass -f python scripts/probe_gemm_cfg.py c4 > $O/r2c_ncu_c4_a.log 2>&1; tail -2 $O/r2c_ncu_c4_a.log
SPX_GEMM_KSPLIT=4 ncu --set full --clock-control none --import-source on -k regex:k_estimate_gemm -s 5 -c 2 -o $O/r2c_gemm_c4_ksplit -f python scripts/probe_gemm_cfg.py c4 > $O/r2c_ncu_c4_b.log 2>&1; tail -2 $O/r2c_ncu_c4_b.log
