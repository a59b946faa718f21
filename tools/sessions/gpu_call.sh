#!/bin/bash
# One GPU-box session: tests, probes, kernel A/B numbers, a bench line.  Output -> gpurun_out/call_<tag>/
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest.txt 2>&1
( timeout 120 tools/probes/copy_floor.bin ) > $OUT/copy_floor.txt 2>&1
for lib in r1 post cipt2 nof2 noh2 copy; do
  echo "== $lib" >> $OUT/kbench.txt
  extra="--no-overlap"; [ $lib = post ] && extra=""
  CERB_LIB=$PWD/cerberusdet_b200/libcerb_$lib.so timeout 300 python tools/kbench.py cfg3 $extra >> $OUT/kbench.txt 2>&1
done
for v in 11 21 22 41 42; do
  echo "== var pipe=$v" >> $OUT/kbench.txt
  CERB_DEBUG_DECODE_PIPE=$v CERB_LIB=$PWD/cerberusdet_b200/libcerb_var.so timeout 300 python tools/kbench.py cfg3 --no-overlap >> $OUT/kbench.txt 2>&1
done
echo "== default, other configs" >> $OUT/kbench.txt
timeout 600 python tools/kbench.py cfg2 cfg4 cfg3f32 cfg3planted >> $OUT/kbench.txt 2>&1
( timeout 600 python bench.py --steps 200 --warmup 10 ) > $OUT/bench.txt 2>&1
tail -3 $OUT/pytest.txt; cat $OUT/kbench.txt | tail -40
