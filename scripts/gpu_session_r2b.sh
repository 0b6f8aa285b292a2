#!/bin/bash
# Round-2 (second half) measurement session on one B200: bench, reference arm, ncu launch list,
# full captures of the hot kernels, the other BASELINE configurations.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r2b_bench.json 2> $O/r2b_bench.err; tail -c 400 $O/r2b_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2b_bench_reference_arm.json 2>> $O/r2b_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2b_launches_bench_steps2_warmup1.csv python bench.py --steps 2 --warmup 1 > $O/r2b_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_dpack|k_sparse_ok|k_estimate_local_fast' -s 6 -c 6 -o $O/r2b_kernels -f python scripts/ncu_target.py 4 > $O/r2b_ncu.log 2>&1; tail -3 $O/r2b_ncu.log
for c in C1 C3 C4 C5; do python bench.py --config $c --steps 3 --warmup 3 > $O/r2b_config_$c.json 2> $O/r2b_config_$c.err; tail -c 300 $O/r2b_config_$c.err; done
ls -la $O | tail -12
