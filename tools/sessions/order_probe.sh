#!/bin/sh
# tools/ only: decode block-order experiment (CERB_DEBUG_DECODE_ORDER = 0 contiguous, 1 interleaved, 2 box-first)
for o in 0 2 1; do
  echo "order $o"
  CERB_DEBUG_DECODE_ORDER=$o timeout 100 python bench.py --steps 400 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['decode_ms_avg'], d['roofline']['frac'])"
  CERB_DEBUG_DECODE_ORDER=$o timeout 60 python tools/microbench.py cfg3f32 2>&1 | tail -1 | cut -c1-110
done
CERB_DEBUG_DECODE_ORDER=2 timeout 100 python -m pytest tests -m gpu -q -k decode 2>&1 | tail -2
