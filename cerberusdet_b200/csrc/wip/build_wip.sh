#!/bin/sh
# Builds the round-2 DRAFT (head_tail.cu) into its own library; nothing in the package loads it.
# Usage: sh cerberusdet_b200/csrc/wip/build_wip.sh   then on a B200: python tools/wip_head_tail_check.py
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
     -o "$HERE/../../libcerb_wip.so" "$HERE/head_tail.cu"
echo "built $HERE/../../libcerb_wip.so"
