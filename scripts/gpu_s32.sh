#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_main.py -m gpu -x -q > $O/s32_tests.log 2>&1; tail -25 $O/s32_tests.log
timeout 300 python __graft_entry__.py smoke > $O/s32_smoke.log 2>&1; tail -3 $O/s32_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/s32_bench.json 2> $O/s32_bench.err; tail -c 1500 $O/s32_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/s32_bench.json').read().strip().splitlines()[-1])
e = d['e2e']
print('value %.3e ms %.3f' % (d['value'], d['ms_per_step']))
print('e2e %.3e ms %.2f B/cell %.3f d2h_gbs %.1f fallbacks %s' % (e['value'], e['ms_per_step'], e['d2h_bytes_per_cell_step'], e['d2h_gbs'], e['fields_sent_through_fallback_codec']))
print('u16', e['u16_transport'], 'raw', e['raw_f32_transport']['value'], 'dec', e['decoded_f32'])
PY
