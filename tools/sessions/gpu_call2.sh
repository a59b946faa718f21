#!/bin/bash
# tests + bench on one GPU.  Output -> gpurun_out/call_<tag>/
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/pytest.txt 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.txt 2>&1
( timeout 900 python bench.py ) > $OUT/bench.txt 2> $OUT/bench.err
( timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline ) > $OUT/bench20.txt 2>> $OUT/bench.err
tail -5 $OUT/pytest.txt; tail -2 $OUT/smoke.txt; tail -c 600 $OUT/bench.err; cut -c1-400 $OUT/bench.txt
