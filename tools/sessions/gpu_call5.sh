#!/bin/bash
# Round-2 evidence session: all GPU tests, smoke, the driver's bench invocations, the ncu launch list of the bench
# command and one --set full capture of each hot kernel.  Output -> gpurun_out/call_<tag>/
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $OUT/pytest.txt 2>&1
( timeout 300 python __graft_entry__.py smoke ) > $OUT/smoke.txt 2>&1
( timeout 600 python bench.py ) > $OUT/bench_default.txt 2> $OUT/bench_default.err
( timeout 600 python bench.py --steps 20 --warmup 5 ) > $OUT/bench20.txt 2> $OUT/bench20.err
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_ref.txt 2> $OUT/bench_ref.err
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras ) > $OUT/bench_under_ncu.txt 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_pipe_kernel -s 6 -c 2 -f -o $OUT/r02_decode \
    python tools/kbench.py cfg3 --no-overlap ) > $OUT/ncu_decode.txt 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_kernel -s 6 -c 2 -f -o $OUT/r02_nms \
    python tools/kbench.py cfg3 --no-overlap ) > $OUT/ncu_nms.txt 2>&1
tail -5 $OUT/pytest.txt; cat $OUT/smoke.txt | tail -2; cut -c1-400 $OUT/bench_default.txt; tail -3 $OUT/bench_default.err
