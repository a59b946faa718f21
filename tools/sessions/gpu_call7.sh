#!/bin/bash
# Head-tail fusion session 2: parity tests, knob A/B (tile order, ring depth), ncu capture, whole GPU suite.
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
( timeout -s KILL 600 python -m pytest tests/test_gpu_headtail.py -q -x -s 2>&1 | tail -40 ) > $OUT/pytest_headtail.txt 2>&1
for v in "1 8" "0 8" "1 6" "1 4" "1 3"; do
  set -- $v
  echo "== ht_order=$1 ht_stages=$2" >> $OUT/headtail_bench.txt
  CERB_DEBUG_HT_ORDER=$1 CERB_DEBUG_HT_STAGES=$2 timeout -s KILL 300 python tools/headtail_bench.py --batch 64 >> $OUT/headtail_bench.txt 2>&1
done
echo "== yolov8n widths" >> $OUT/headtail_bench.txt
timeout -s KILL 300 python tools/headtail_bench.py --batch 64 --c2 64 --c3 64 >> $OUT/headtail_bench.txt 2>&1
( timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:head_tail_kernel -s 3 -c 1 -f -o $OUT/r02_head_tail \
    python tools/headtail_bench.py --batch 64 --reps 2 ) > $OUT/ncu_headtail.txt 2>&1
( timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $OUT/pytest_all.txt 2>&1
tail -12 $OUT/pytest_headtail.txt; cat $OUT/headtail_bench.txt | cut -c1-700; tail -6 $OUT/pytest_all.txt
