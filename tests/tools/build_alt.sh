#!/bin/bash
# Build the kernels of git revision $1 (default HEAD) into crick_b200/libcrick_cuda_alt.so for
# same-box A/B runs: CRICK_CUDA_LIB=$PWD/crick_b200/libcrick_cuda_alt.so python bench.py ...
set -e
rev=${1:-HEAD}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
git -C "$root" archive "$rev" crick_b200/csrc include | tar -x -C "$tmp"
make -C "$tmp/crick_b200/csrc" -j8 > /dev/null
cp "$tmp/crick_b200/libcrick_cuda.so" "$root/crick_b200/libcrick_cuda_alt.so"
rm -rf "$tmp"
echo built "$root/crick_b200/libcrick_cuda_alt.so" from "$rev"
