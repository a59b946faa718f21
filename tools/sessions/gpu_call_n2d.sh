#!/bin/bash
# N-GPU diagnostic: independent replicas / peer delivery variants / gather, all on the same box (box-to-box variation is +-5 %)
TAG=${1:-x}
N=${2:-2}
OUT=gpurun_out/call_$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581"
( timeout -s KILL 400 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -30 ) > $OUT/pytest_multi.txt 2>&1
run() { name=$1; shift; ( env "$@" timeout -s KILL 400 $TR bench.py --gpus $N --steps 400 --warmup 20 --no-cpu-baseline --no-extras ) > $OUT/bench_$name.txt 2> $OUT/bench_$name.err; }
run none CERB_DELIVERY=none
run peer CERB_DELIVERY=peer
run gather CERB_DELIVERY=gather
run none_streams CERB_DELIVERY=none CERB_SCHEDULE=streams
run peer_last3 CERB_DELIVERY=peer CERB_SIDE=last3 CERB_SCHEDULE=streams
run peer2 CERB_DELIVERY=peer
( env CERB_DELIVERY=peer timeout -s KILL 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extras ) > $OUT/bench_peer20.txt 2> $OUT/bench_peer20.err
( timeout -s KILL 300 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-extras ) > $OUT/bench_n1.txt 2> $OUT/bench_n1.err
tail -3 $OUT/pytest_multi.txt
for f in n1 none peer gather none_streams peer_last3 peer2 peer20; do echo $f; grep -o '"ms_per_step": [0-9.]*\|"per_rank_ms": {[^}]*}\|"n_gpu_equals_1_gpu": [a-z]*' $OUT/bench_$f.txt | head -3; tail -c 600 $OUT/bench_$f.err | grep -i "error\|Traceback" ; done
