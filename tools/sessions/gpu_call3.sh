#!/bin/bash
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $OUT/pytest.txt 2>&1
( timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline ) > $OUT/bench20.txt 2> $OUT/bench.err
( timeout 300 python bench.py --steps 400 --warmup 20 --no-extras --no-cpu-baseline ) > $OUT/bench400.txt 2>> $OUT/bench.err
CERB_LIB=$PWD/cerberusdet_b200/libcerb_prof.so timeout 300 python tools/nms_phases.py cfg3 cfg3planted cfg2 cfg4 > $OUT/phases.txt 2>&1
cp gpurun_out/torchvision_cuda_threshold.json $OUT/ 2>/dev/null
tail -12 $OUT/pytest.txt; tail -c 600 $OUT/bench.err; cat $OUT/phases.txt
