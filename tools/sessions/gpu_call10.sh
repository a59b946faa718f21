#!/bin/bash
# Validation session: all GPU tests, smoke, the driver's bench invocations, launch list, head-tail ncu capture.
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
( timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $OUT/pytest.txt 2>&1
( timeout -s KILL 300 python __graft_entry__.py smoke ) > $OUT/smoke.txt 2>&1
( timeout -s KILL 600 python bench.py ) > $OUT/bench_default.txt 2> $OUT/bench_default.err
( timeout -s KILL 600 python bench.py --steps 20 --warmup 5 ) > $OUT/bench20.txt 2> $OUT/bench20.err
( timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras ) > $OUT/bench_under_ncu.txt 2>&1
( timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:head_tail_kernel -s 3 -c 1 -f -o $OUT/r02_head_tail \
    python tools/headtail_bench.py --batch 64 --reps 2 ) > $OUT/ncu_headtail.txt 2>&1
tail -5 $OUT/pytest.txt; tail -2 $OUT/smoke.txt; cut -c1-300 $OUT/bench_default.txt; tail -3 $OUT/bench_default.err
