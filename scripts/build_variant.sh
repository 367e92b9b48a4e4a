#!/bin/bash
# Development helper: build an A/B variant of the library with extra -D flags.
#   scripts/build_variant.sh <suffix> [-DNAME=VALUE ...]   ->  cnmf_e_b200/libcnmfe_b200_<suffix>.so   (use with CNMFE_B200_LIB=...)
set -e
cd "$(dirname "$0")/.."
suf=$1; shift
out=cnmf_e_b200/build/var_$suf
mkdir -p $out
for f in cnmf_e_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --fmad=false "$@" -c $f -o $out/$(basename $f).o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o cnmf_e_b200/libcnmfe_b200_$suf.so $out/*.o -lcudart
echo built cnmf_e_b200/libcnmfe_b200_$suf.so
