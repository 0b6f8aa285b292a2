#!/bin/bash
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s39_tests.log 2>&1; tail -15 $O/s39_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/s39_bench.json 2> $O/s39_bench.err; tail -c 1500 $O/s39_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/s39_bench.json').read().strip().splitlines()[-1])
e = d['e2e']
print('value %.3e ms %.3f' % (d['value'], d['ms_per_step']), 'roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'])
print('e2e %.3e ms %.2f B/cell %.3f d2h_gbs %.1f fallbacks %s' % (e['value'], e['ms_per_step'], e['d2h_bytes_per_cell_step'], e['d2h_gbs'], e['fields_sent_through_fallback_codec']))
print('breakdown', d['step_breakdown'])
print('dense', d['dense_path']['value'], d['dense_path']['ms_per_step'])
PY
