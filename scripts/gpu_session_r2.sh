#!/bin/bash
# Round-2 measurement session on one B200: tests, bench, ncu launch list + full captures,
# the other BASELINE configurations.  Everything lands in gpurun_out/.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2_tests.log 2>&1; tail -4 $O/r2_tests.log
python bench.py --steps 20 --warmup 5 > $O/r2_bench.json 2> $O/r2_bench.err; tail -c 400 $O/r2_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference_arm.json 2>> $O/r2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench_steps2_warmup1.csv python bench.py --steps 2 --warmup 1 > $O/r2_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_estimate_local_fast|k_downdate_lu|k_ut_gemm|k_pack_scan|k_pack_encode|k_round_stats' -s 7 -c 14 -o $O/r2_kernels -f python scripts/ncu_target.py 4 > $O/r2_ncu.log 2>&1; tail -3 $O/r2_ncu.log
for c in C1 C3 C4 C5; do python bench.py --config $c --steps 3 --warmup 3 > $O/r2_config_$c.json 2> $O/r2_config_$c.err; tail -c 300 $O/r2_config_$c.err; done
ls -la $O | tail -15
