#!/bin/bash
# Development helper: builds a variant of libawfm_b200.so with extra -D flags into
# avxwindowfmindex_b200/csrc/variants/<name>/ (git-ignored, travels to the GPU box) for A/B probes:
#   tools/build_variant.sh two_sectors -DAWFM_SWEEP_TWO_SECTORS
#   AWFM_B200_LIB=$PWD/avxwindowfmindex_b200/csrc/variants/two_sectors/libawfm_b200.so python tools/sweep_probe.py ...
# SOURCES="awfm_b200 awfm_multi" (default: awfm_b200) names the translation units the flags apply to.
set -e
name=$1; shift
cd "$(dirname "$0")/../avxwindowfmindex_b200/csrc"
mkdir -p variants/$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
objs=""
for tu in awfm_b200 awfm_build awfm_multi; do
  if [[ " ${SOURCES:-awfm_b200} " == *" $tu "* ]]; then
    /usr/local/cuda/bin/nvcc $ARCH "$@" -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fopenmp,-Wno-deprecated-declarations -Wno-deprecated-declarations -c -o variants/$name/$tu.o $tu.cu &
    objs="$objs variants/$name/$tu.o"
  else
    objs="$objs $tu.o"
  fi
done
wait
/usr/local/cuda/bin/nvcc $ARCH -shared -Xcompiler -fPIC -o variants/$name/libawfm_b200.so $objs awfm_dropin.o -lgomp -lpthread
echo built variants/$name/libawfm_b200.so
