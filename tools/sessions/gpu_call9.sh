#!/bin/bash
# Head-tail fusion session 4: adaptive epilogue groups + backoff; groups knob A/B at three widths.
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
( timeout -s KILL 600 python -m pytest tests/test_gpu_headtail.py -q -x 2>&1 | tail -15 ) > $OUT/pytest_headtail.txt 2>&1
for w in "80 320" "64 256" "64 192" "64 128" "64 64"; do
  set -- $w
  for g in 1 2; do
    echo "== c2=$1 c3=$2 groups=$g" >> $OUT/headtail_bench.txt
    CERB_DEBUG_HT_GROUPS=$g timeout -s KILL 300 python tools/headtail_bench.py --batch 64 --c2 $1 --c3 $2 --reps 10 >> $OUT/headtail_bench.txt 2>&1
  done
done
echo "== c2=80 c3=320 groups=1 order=0" >> $OUT/headtail_bench.txt
CERB_DEBUG_HT_GROUPS=1 CERB_DEBUG_HT_ORDER=0 timeout -s KILL 300 python tools/headtail_bench.py --batch 64 --reps 10 >> $OUT/headtail_bench.txt 2>&1
tail -4 $OUT/pytest_headtail.txt; grep -o '== .*\|"fused_us": [0-9.]*\|"fused_GBps": [0-9.]*\|"unfused_reference_order_us": [0-9.]*' $OUT/headtail_bench.txt | paste - - - -
