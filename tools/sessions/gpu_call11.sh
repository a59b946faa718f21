#!/bin/bash
# Single-stream PDL overlapped schedule vs two graph branches, same box; then the whole suite and the bench lines.
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
for s in pdl streams pdl streams; do
  echo "== schedule=$s" >> $OUT/sched.txt
  CERB_SCHEDULE=$s timeout -s KILL 120 python tools/side_probe.py --steps 400 --one none >> $OUT/sched.txt 2>&1
done
for s in pdl streams; do
  ( CERB_SCHEDULE=$s timeout -s KILL 300 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-extras ) > $OUT/bench_$s.txt 2> $OUT/bench_$s.err
  ( CERB_SCHEDULE=$s timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras ) > $OUT/bench20_$s.txt 2>> $OUT/bench_$s.err
done
( timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $OUT/pytest.txt 2>&1
cat $OUT/sched.txt | cut -c1-200; for f in bench_pdl bench_streams bench20_pdl bench20_streams; do echo $f; grep -o '"ms_per_step": [0-9.]*\|"decode_ms_avg": [0-9.]*\|"nms_ms_avg": [0-9.]*\|"pipeline_equals_direct_calls": [a-z]*' $OUT/$f.txt | tr '\n' ' '; echo; done; tail -4 $OUT/pytest.txt
