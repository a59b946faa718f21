#!/bin/bash
# 2-GPU session: sharded == single-GPU through both deliveries, bench at N=2 with each.
TAG=${1:-x}
N=${2:-2}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
( timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -30 ) > $OUT/pytest_multi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581"
( timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline ) > $OUT/bench_peer.txt 2> $OUT/bench_peer.err
( CERB_DELIVERY=gather timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --no-extras ) > $OUT/bench_gather.txt 2> $OUT/bench_gather.err
( timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extras ) > $OUT/bench_peer20.txt 2>> $OUT/bench_peer.err
tail -6 $OUT/pytest_multi.txt; tail -c 1500 $OUT/bench_peer.err; cut -c1-300 $OUT/bench_peer.txt; cut -c1-300 $OUT/bench_gather.txt
