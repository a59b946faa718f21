#!/bin/sh
# Build libcerb_post.so in-tree for sm_100a.  Usage: sh cerberusdet_b200/csrc/build.sh [extra nvcc flags]
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="${CERB_OUT:-$HERE/../libcerb_post.so}"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC -shared "$@" \
     -o "$OUT" "$HERE/decode.cu" "$HERE/decode_tma.cu" "$HERE/decode_pipe.cu" "$HERE/nms.cu" "$HERE/cross_task.cu" "$HERE/val_match.cu" "$HERE/train_decode.cu" "$HERE/api.cu"
echo "built $OUT"
