#!/bin/bash
# 8-GPU check of the shipped configuration (single-stream PDL schedule + piggyback delivery) beside N independent replicas
TAG=${1:-x}
N=${2:-8}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581"
run() { name=$1; steps=$2; warm=$3; shift 3; ( env "$@" timeout -s KILL 200 $TR bench.py --gpus $N --steps $steps --warmup $warm --no-cpu-baseline --no-extras ) > $OUT/bench_$name.txt 2> $OUT/bench_$name.err; }
run none 400 20 CERB_DELIVERY=none
run peer 400 20 CERB_DELIVERY=peer
run peer20 20 5 CERB_DELIVERY=peer
for f in none peer peer20; do echo $f; grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.]*, "unit": "images/s", "n_gpus"\|"elapsed": \[[^]]*\]\|"n_gpu_equals_1_gpu": [a-z]*' $OUT/bench_$f.txt | head -4; tail -c 600 $OUT/bench_$f.err | grep -i "error\|Traceback" ; done
