#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r2c_two_gpu_tests.log 2>&1; tail -3 $O/r2c_two_gpu_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2c_bench_n2.json 2> $O/r2c_bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2c_bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
e = d['e2e']
print('N=2 value %.4e ms %.4f' % (d['value'], d['ms_per_step']), 'e2e %.4e ms %.2f d2h_gbs %.1f' % (e['value'], e['ms_per_step'], e['d2h_gbs']), 'u16 ms', e['u16_transport']['ms_per_step'], 'gather', d['gather']['gbs_into_writer'])
PY
