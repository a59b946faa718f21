#!/bin/bash
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
( timeout -s KILL 600 python -m pytest tests/test_gpu_tal.py -q 2>&1 | tail -10 ) > $OUT/pytest_tal.txt 2>&1
( timeout -s KILL 300 python tools/microbench_tal.py ) > $OUT/tal_bench.txt 2>&1
( timeout -s KILL 300 python tools/microbench_tal.py --batch 16 --boxes 60 --classes 80 ) >> $OUT/tal_bench.txt 2>&1
( timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tal_ -c 12 --csv --log-file $OUT/tal_launches.csv python tools/microbench_tal.py ) > $OUT/ncu_tal.txt 2>&1
tail -4 $OUT/pytest_tal.txt; tail -3 $OUT/tal_bench.txt; grep -o '"tal_[a-z_]*kernel[^"]*".*' $OUT/tal_launches.csv | awk -F'"' '{print $1 $2, $(NF-1)}' | tail -6
