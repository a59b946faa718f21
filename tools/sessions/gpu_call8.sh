#!/bin/bash
# Head-tail fusion session 3: two epilogue groups.  Parity, timing at two widths, ncu capture.
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
( timeout -s KILL 600 python -m pytest tests/test_gpu_headtail.py -q -x -s 2>&1 | tail -40 ) > $OUT/pytest_headtail.txt 2>&1
echo "== yolov8x widths" >> $OUT/headtail_bench.txt
timeout -s KILL 300 python tools/headtail_bench.py --batch 64 >> $OUT/headtail_bench.txt 2>&1
echo "== yolov8n widths" >> $OUT/headtail_bench.txt
timeout -s KILL 300 python tools/headtail_bench.py --batch 64 --c2 64 --c3 64 >> $OUT/headtail_bench.txt 2>&1
echo "== yolov8l widths" >> $OUT/headtail_bench.txt
timeout -s KILL 300 python tools/headtail_bench.py --batch 64 --c2 64 --c3 256 >> $OUT/headtail_bench.txt 2>&1
( timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:head_tail_kernel -s 3 -c 1 -f -o $OUT/r02_head_tail \
    python tools/headtail_bench.py --batch 64 --reps 2 ) > $OUT/ncu_headtail.txt 2>&1
tail -8 $OUT/pytest_headtail.txt; cat $OUT/headtail_bench.txt | cut -c1-700
