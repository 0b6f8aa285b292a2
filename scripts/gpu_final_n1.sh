#!/bin/bash
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c_tests.log 2>&1; tail -4 $O/r2c_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python scripts/probe_first_chunk.py 2>&1 | grep "chunk"
timeout 300 python scripts/probe_job.py 2>&1 | grep "^rep"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2c_bench.json 2> $O/r2c_bench.err; tail -c 300 $O/r2c_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value %.4e ms %.4f solve %.4f est %.4f frac %.4f' % (d['value'], d['ms_per_step'], d['solve_phase']['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac']))
print('e2e %.4e ms %.3f dec %.3e u16 %.3e raw %.3e' % (e['value'], e['ms_per_step'], e['decoded_f32']['value'], e['u16_transport']['value'], e['raw_f32_transport']['value']))
print(d['clocks'])
"
