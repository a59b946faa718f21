#!/bin/bash
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > $OUT/pytest.txt 2>&1
for lib in post minb6 minb4 t64; do
  echo "== $lib" >> $OUT/kbench.txt
  CERB_LIB=$PWD/cerberusdet_b200/libcerb_$lib.so timeout 300 python tools/kbench.py cfg3 --no-overlap >> $OUT/kbench.txt 2>&1
done
tail -8 $OUT/pytest.txt; cat $OUT/kbench.txt
