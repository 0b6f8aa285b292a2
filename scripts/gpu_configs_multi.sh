#!/bin/bash
# bench.py --config at N GPUs (argument 1), every configuration; outputs in gpurun_out/
N=$1
O=gpurun_out
for c in C1 C3 C4 C5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --config $c --steps 3 --warmup 3 > $O/r2_config_${c}_n$N.json 2> $O/r2_config_${c}_n$N.err
  tail -c 300 $O/r2_config_${c}_n$N.err | grep -v Warn
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/r2_config_${c}_n$N.json") if l.startswith("{")][-1])
    print("$c N=$N", "%.3e" % d["value"], "ms/step", round(d["ms_per_step"],2), d["roofline"] and (d["roofline"]["kernel"], round(d["roofline"]["frac"],3)))
except Exception as e:
    print("$c N=$N failed", e)
PY
done
