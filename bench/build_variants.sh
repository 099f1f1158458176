#!/bin/bash
# builds experimental variants of libb200iop.so into bench/variants/ (kernel tuning experiments)
set -e
cd "$(dirname "$0")/../libiop_b200/csrc"
OUT=../../bench/variants
mkdir -p $OUT
build() { # name flags
  name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr "$@" \
    -shared -o $OUT/lib_$name.so api.cu additive_fft.cu ntt.cu fri.cu merkle.cu elementwise.cu -cudart static -Xcompiler -nostdlib++ -Xlinker /usr/lib/x86_64-linux-gnu/libstdc++.so.6 &
}
for v in "$@"; do
  case $v in
    t256)   build t256   -DB200_TILE_THREADS=256 -DB200_TILE_MINBLOCKS=2 -DB200_ILP=1 ;;
    t512)   build t512   -DB200_TILE_THREADS=512 -DB200_TILE_MINBLOCKS=2 -DB200_ILP=1 ;;
    t256i2) build t256i2 -DB200_TILE_THREADS=256 -DB200_TILE_MINBLOCKS=2 -DB200_ILP=2 ;;
    t512i2) build t512i2 -DB200_TILE_THREADS=512 -DB200_TILE_MINBLOCKS=2 -DB200_ILP=2 ;;
    notw)   build notw   -DB200_EXP_NOTW ;;
    nosync) build nosync -DB200_EXP_NOSYNC ;;
    nomul)  build nomul  -DB200_EXP_NOMUL ;;
    base)   build base ;;
    s4)     build s4     -DB200_GF_SPACING=4 ;;
    s5)     build s5     -DB200_GF_SPACING=5 ;;
    t384)   build t384   -DB200_TILE_THREADS=384 -DB200_TILE_MINBLOCKS=2 -DB200_ILP=1 ;;
  esac
done
wait
ls -la $OUT
