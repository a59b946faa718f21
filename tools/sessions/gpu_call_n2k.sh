#!/bin/bash
# the driver's invocations at N = 1 and N = 2 on one box, with the steady-state field
TAG=${1:-x}
N=${2:-2}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581"
( timeout -s KILL 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras ) > $OUT/bench_n1_20.txt 2> $OUT/bench_n1_20.err
( timeout -s KILL 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extras ) > $OUT/bench_n${N}_20.txt 2> $OUT/bench_n${N}_20.err
( timeout -s KILL 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline ) > $OUT/bench_n${N}_20_extras.txt 2> $OUT/bench_n${N}_20_extras.err
( timeout -s KILL 300 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -5 ) > $OUT/pytest_multi.txt 2>&1
for f in bench_n1_20 bench_n${N}_20 bench_n${N}_20_extras; do echo $f; grep -o '"ms_per_step": [0-9.]*\|"steady_state": {[^}]*}\|"n_gpu_equals_1_gpu": [a-z]*\|"config5_strong": {[^}]*}' $OUT/$f.txt | head -5; tail -c 500 $OUT/$f.err | grep -i "error\|Traceback"; done; tail -2 $OUT/pytest_multi.txt
