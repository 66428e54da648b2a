#!/bin/bash
# Developer A/B builds: scripts/build_variant.sh NAME [extra nvcc flags...] -> ab/NAME.so (git-ignored, travels with gpurun)
set -euo pipefail
NAME=$1; shift
cd "$(dirname "$0")/../prophylo_b200/csrc"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH $*"
B=build_$NAME; mkdir -p $B
nvcc $COMMON -c ppp_sweep.cu -o $B/ppp_sweep.o &
nvcc $COMMON -fmad=false -c ppp_exact.cu -o $B/ppp_exact.o &
nvcc $COMMON -c ppp_topk.cu -o $B/ppp_topk.o &
nvcc $COMMON -c ppp_ranked.cu -o $B/ppp_ranked.o &
nvcc $COMMON -c ppp_prefilter.cu -o $B/ppp_prefilter.o &
nvcc $COMMON -c ppp_prefilter_tc.cu -o $B/ppp_prefilter_tc.o &
nvcc $COMMON -c ppp_debug.cu -o $B/ppp_debug.o &
nvcc $COMMON -fmad=false -c ppp_table.cu -o $B/ppp_table.o &
nvcc $COMMON -c ppp_rank.cu -o $B/ppp_rank.o &
nvcc $COMMON -c ppp_capi.cu -o $B/ppp_capi.o &
g++ -O2 -std=c++17 -fPIC -pthread -c ppp_bla.cpp -o $B/ppp_bla.o &
g++ -O2 -std=c++17 -fPIC -pthread -c ppp_pack.cpp -o $B/ppp_pack.o &
FAIL=0; for j in $(jobs -p); do wait $j || FAIL=1; done; [ $FAIL = 0 ] || { echo "nvcc failed"; exit 1; }
mkdir -p ../../ab
nvcc $ARCH -shared -o ../../ab/$NAME.so $B/*.o
rm -rf $B
echo "built ab/$NAME.so"
