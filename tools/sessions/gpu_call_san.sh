#!/bin/bash
# compute-sanitizer over the kernels added in round 2
TAG=${1:-x}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
( timeout -s KILL 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_tal.py -q -x -k "golden or oracle_port" 2>&1 | tail -12 ) > $OUT/memcheck_tal.txt 2>&1
( timeout -s KILL 500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_tal.py -q -x -k "golden" 2>&1 | tail -12 ) > $OUT/racecheck_tal.txt 2>&1
( timeout -s KILL 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_headtail.py -q -x -k "golden or oracle_port or rejects" 2>&1 | tail -12 ) > $OUT/memcheck_headtail.txt 2>&1
( timeout -s KILL 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -q -x -k "pipeline_equals_direct_calls" 2>&1 | tail -12 ) > $OUT/memcheck_pipeline.txt 2>&1
( timeout -s KILL 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_reference.py -q -x -k "validation_batch_statistics" 2>&1 | tail -12 ) > $OUT/memcheck_val.txt 2>&1
for f in memcheck_tal racecheck_tal memcheck_headtail memcheck_pipeline memcheck_val; do echo "== $f"; grep -E "passed|failed|ERROR SUMMARY|error" $OUT/$f.txt | tail -3; done
