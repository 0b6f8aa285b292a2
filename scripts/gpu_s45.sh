#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_main.py -m gpu -x -q > $O/s45_tests.log 2>&1; tail -4 $O/s45_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python scripts/run_configs.py C2 > $O/s45_c2_full.log 2>&1; tail -3 $O/s45_c2_full.log | cut -c1-1500
SPX_SPARSE_SOLVE=0 timeout 600 python scripts/run_configs.py C2 > $O/s45_c2_full_nosparse.log 2>&1; tail -1 $O/s45_c2_full_nosparse.log | cut -c1-600
timeout 600 python bench.py --steps 20 --warmup 5 > $O/s45_bench.json 2> $O/s45_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/s45_bench.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value %.4e ms %.4f solve %.4f est %.4f' % (d['value'], d['ms_per_step'], d['solve_phase']['ms_per_step'], d['roofline']['avg_launch_ms']))
print('e2e %.4e ms %.3f dec %.3e u16 %.3e' % (e['value'], e['ms_per_step'], e['decoded_f32']['value'], e['u16_transport']['value']))
"
