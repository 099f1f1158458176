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
    vec64)  build vec64  -DB200_TILE_VEC128=0 ;;          # round 2: 64-bit tile loads / stores (the round-1 kernel)
    vec128) build vec128 -DB200_TILE_VEC128=1 ;;
    blake0) build blake0 -DB200_BLAKE_IMAD=0 ;;           # round 2: BLAKE2b additions on the ALU pipe (round-1 kernel)
    blake1) build blake1 -DB200_BLAKE_IMAD=1 ;;           #          c += d on the FMA pipe
    blake2) build blake2 -DB200_BLAKE_IMAD=2 ;;           #          all additions on the FMA pipe
    ntt11)  build ntt11  -DB200_NTT_TILE_BITS=11 ;;       # round 2: NTT tiles of 2^11 elements (4 CTAs per SM)
    ntt12)  build ntt12  -DB200_NTT_TILE_BITS=12 ;;
    ntt10)  build ntt10  -DB200_NTT_TILE_BITS=10 ;;
    leaf2)  build leaf2  -DB200_LEAF1_MINBLOCKS=2 ;;      # round 2: register budget of merkle_leaf1_kernel (2 / 3 / 4 CTAs of 256 per SM)
    leaf3)  build leaf3  -DB200_LEAF1_MINBLOCKS=3 ;;
    leaf4)  build leaf4  -DB200_LEAF1_MINBLOCKS=4 ;;
    inv8)   build inv8   -DB200_INV_RUN=8 ;;              # round 2: elements per Itoh-Tsujii inversion in gf192_inv_kernel
    inv16)  build inv16  -DB200_INV_RUN=16 ;;
    inv32)  build inv32  -DB200_INV_RUN=32 ;;
  esac
done
wait
ls -la $OUT
