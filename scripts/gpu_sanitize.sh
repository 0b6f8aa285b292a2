#!/bin/bash
# compute-sanitizer over the kernels added in the second half of round 2
O=gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_main.py -m gpu -x -q -k "delta or fused or packed" > $O/san_memcheck_pack.log 2>&1; echo "memcheck pack rc=$?"; tail -4 $O/san_memcheck_pack.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "sparse" > $O/san_memcheck_sparse.log 2>&1; echo "memcheck sparse rc=$?"; tail -4 $O/san_memcheck_sparse.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_main.py -m gpu -x -q -k "delta_encoder or fused" > $O/san_racecheck_pack.log 2>&1; echo "racecheck pack rc=$?"; tail -4 $O/san_racecheck_pack.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_main.py -m gpu -x -q -k "delta_encoder or fused" > $O/san_synccheck_pack.log 2>&1; echo "synccheck pack rc=$?"; tail -4 $O/san_synccheck_pack.log
