#!/bin/bash
# Builds variants of the CTA-pair matcher (csrc/hamming_tc2.cu) with different compile-time knobs into
# tools/tc2_variants/ (git-ignored; they ride to the GPU box with the snapshot) -- run this where nvcc is.
# Then, on the GPU:  for f in tools/tc2_variants/*.so; do TVG_B200_LIB=$PWD/$f python tools/knn_only.py; done
set -e
cd "$(dirname "$0")/.."
OUT=tools/tc2_variants; mkdir -p $OUT
OBJ=openwork_b200/_lib
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
build() {  # name, -D flags
  nvcc $FLAGS $2 -c openwork_b200/csrc/hamming_tc2.cu -o $OUT/$1.o
  objs=$(ls $OBJ/*.o | grep -v hamming_tc2.o)
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libtvg_$1.so $objs $OUT/$1.o -ldl
  rm $OUT/$1.o
}
if [ "$1" = "epilogue" ]; then   # how the epilogue overlaps its tensor-memory loads with the arithmetic
  build spl0 "-DT2_SPLIT_LD_N=0"
  build spl1 "-DT2_SPLIT_LD_N=1"
  build spl2 "-DT2_SPLIT_LD_N=2"
  build halves "-DT2_HALVES_N=1"
  build first2 "-DT2_SPLIT_LD_N=1 -DT2_FIRST_LD_N=2"
  build early0 "-DT2_EARLY_LD_N=0"
  build early1 "-DT2_EARLY_LD_N=1"
else
  build base ""
  build lists2 "-DT2_LISTS=2"
  build lists8 "-DT2_LISTS=8"
  build exp2 "-DT2_EXP_WARPS_N=2"
  build exp4 "-DT2_EXP_WARPS_N=4"
  build nst3 "-DT2_NST_N=3"
fi
ls -la $OUT
