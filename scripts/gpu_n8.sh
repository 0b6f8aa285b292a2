#!/bin/bash
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2b_bench_n8.json 2> $O/r2b_bench_n8.err; tail -c 300 $O/r2b_bench_n8.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2b_bench_n8.json').read().strip().splitlines() if l.startswith('{')][-1])
e = d['e2e']
print('N=8 value %.4e ms %.4f wall %.4f solve %s' % (d['value'], d['ms_per_step'], d['config']['wall_ms_per_step'], d['solve_phase']['ms_per_step']))
print('e2e %.4e ms %.2f d2h_gbs/rank %.1f B/cell %.3f' % (e['value'], e['ms_per_step'], e['d2h_gbs'], e['d2h_bytes_per_cell_step']))
print('u16', e['u16_transport']['ms_per_step'], 'raw', e['raw_f32_transport']['ms_per_step'], 'dec', e['decoded_f32']['ms_per_step'])
print('gather', d['gather'])
PY
nproc; free -g | head -2
