#!/bin/bash
# Head-tail fusion session: first run of the tcgen05 kernel (draft check, parity tests, fused-vs-unfused timing).
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
( timeout -s KILL 120 python tools/wip_head_tail_check.py ) > $OUT/wip_check.txt 2>&1
( timeout -s KILL 600 python -m pytest tests/test_gpu_headtail.py -q -x 2>&1 | tail -40 ) > $OUT/pytest_headtail.txt 2>&1
( timeout -s KILL 300 python tools/headtail_bench.py --batch 64 ) > $OUT/headtail_bench.txt 2>&1
( timeout -s KILL 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_headtail.py -q -x -k "golden or rejects" 2>&1 | tail -15 ) > $OUT/sanitizer_headtail.txt 2>&1
cat $OUT/wip_check.txt | tail -8; tail -25 $OUT/pytest_headtail.txt; tail -5 $OUT/headtail_bench.txt; tail -5 $OUT/sanitizer_headtail.txt
