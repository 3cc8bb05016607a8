#!/bin/bash
# Build lib/libctts_b200_ab.so from the csrc/ of a git revision (default HEAD) for tools/ab_lib.sh: A/B of a kernel change on one box.
set -e
REV=${1:-HEAD}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
cd "$ROOT"
for f in $(git ls-files comprehensive-transformer-tts_b200/csrc); do git show $REV:$f > $TMP/$(basename $f); done
git show $REV:include/ctts_b200.h > $TMP/ctts_b200.h
cd $TMP
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -I . \
    ctts_ops.cu ctts_gemm_tc.cu ctts_flash_attn.cu ctts_train.cu -o "$ROOT/comprehensive-transformer-tts_b200/lib/libctts_b200_ab.so"
rm -rf $TMP
