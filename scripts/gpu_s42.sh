#!/bin/bash
O=gpurun_out
nvidia-smi topo -m > $O/s42_topo.log 2>&1; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" >> $O/s42_topo.log; python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))" >> $O/s42_topo.log; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/class); fi; done >> $O/s42_topo.log; free -g >> $O/s42_topo.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/s42_tests.log 2>&1; tail -5 $O/s42_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/s42_bench_n2.json 2> $O/s42_bench_n2.err; tail -c 600 $O/s42_bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/s42_bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
e = d['e2e']
print('N=2 value %.3e ms %.3f' % (d['value'], d['ms_per_step']), 'e2e %.3e ms %.2f d2h_gbs %.1f' % (e['value'], e['ms_per_step'], e['d2h_gbs']), 'gather', d['gather'])
PY
cat $O/s42_topo.log | head -40
