#!/bin/bash
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 scripts/value_loop.py 40 > $O/n8_$tag.log 2>&1; echo "== $tag"; grep "rep 2" $O/n8_$tag.log | sort | head -8; }
run default A=1
run threads1 SPX_HOST_THREADS=1
run nodist NO_DIST=1
run nodist_t1 NO_DIST=1 SPX_HOST_THREADS=1
