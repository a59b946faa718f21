#!/bin/bash
# 2-GPU session: in-kernel delivery.  Sharded == single-GPU through both deliveries, bench at N with each.
TAG=${1:-x}
N=${2:-2}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
( timeout -s KILL 400 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -30 ) > $OUT/pytest_multi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581"
( timeout -s KILL 400 $TR bench.py --gpus $N --steps 400 --warmup 20 --no-cpu-baseline ) > $OUT/bench_peer.txt 2> $OUT/bench_peer.err
( timeout -s KILL 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extras ) > $OUT/bench_peer20.txt 2>> $OUT/bench_peer.err
( CERB_DELIVERY=gather timeout -s KILL 400 $TR bench.py --gpus $N --steps 400 --warmup 20 --no-cpu-baseline --no-extras ) > $OUT/bench_gather.txt 2> $OUT/bench_gather.err
( timeout -s KILL 300 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-extras ) > $OUT/bench_n1.txt 2> $OUT/bench_n1.err
tail -6 $OUT/pytest_multi.txt; tail -c 1200 $OUT/bench_peer.err; for f in bench_n1 bench_peer bench_peer20 bench_gather; do echo $f; grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus": [0-9]*\|"ms_per_step": [0-9.]*\|"n_gpu_equals_1_gpu": [a-z]*\|"config5_strong": {[^}]*}' $OUT/$f.txt | head -6; done
