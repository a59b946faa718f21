#!/bin/bash
# TAL assigner + validation-statistics session: the new GPU tests, a timing, then the whole suite.
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
( timeout -s KILL 600 python -m pytest tests/test_gpu_tal.py -q 2>&1 | tail -40 ) > $OUT/pytest_tal.txt 2>&1
( timeout -s KILL 300 python -m pytest tests/test_gpu_reference.py -q -k "validation_batch_statistics" 2>&1 | tail -30 ) > $OUT/pytest_val.txt 2>&1
( timeout -s KILL 300 python tools/microbench_tal.py ) > $OUT/tal_bench.txt 2>&1
( timeout -s KILL 300 python tools/microbench_tal.py --batch 16 --boxes 60 --classes 80 ) >> $OUT/tal_bench.txt 2>&1
( timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $OUT/pytest_all.txt 2>&1
tail -25 $OUT/pytest_tal.txt; tail -15 $OUT/pytest_val.txt; tail -3 $OUT/tal_bench.txt; tail -5 $OUT/pytest_all.txt
