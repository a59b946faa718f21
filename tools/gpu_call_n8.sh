#!/bin/bash
# 8-GPU A/B of the deliveries on one box (each run ~20 s)
TAG=${1:-x}
N=${2:-8}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581"
run() { name=$1; shift; ( env "$@" timeout -s KILL 200 $TR bench.py --gpus $N --steps 400 --warmup 20 --no-cpu-baseline --no-extras ) > $OUT/bench_$name.txt 2> $OUT/bench_$name.err; }
run none CERB_DELIVERY=none
run branch3 CERB_DELIVERY=peer CERB_SIDE=branch3
run gather CERB_DELIVERY=gather
run piggy CERB_DELIVERY=peer
( env CERB_DELIVERY=peer CERB_SIDE=branch3 timeout -s KILL 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extras ) > $OUT/bench_branch3_20.txt 2> $OUT/bench_branch3_20.err
( env CERB_DELIVERY=gather timeout -s KILL 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extras ) > $OUT/bench_gather_20.txt 2> $OUT/bench_gather_20.err
for f in none branch3 gather piggy branch3_20 gather_20; do echo $f; grep -o '"ms_per_step": [0-9.]*\|"elapsed": \[[^]]*\]\|"n_gpu_equals_1_gpu": [a-z]*' $OUT/bench_$f.txt | head -3; tail -c 600 $OUT/bench_$f.err | grep -i "error\|Traceback" ; done
