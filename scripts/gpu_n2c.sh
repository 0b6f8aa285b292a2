#!/bin/bash
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 scripts/value_loop.py 200 > $O/n2_$tag.log 2>&1; echo "== $tag"; grep "rep [12]" $O/n2_$tag.log | sort | cut -c1-60; }
run nosampler A=1
run sampler VL_SAMPLER=1
run sampler_prof VL_SAMPLER=1 VL_PROF=1
