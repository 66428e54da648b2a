#!/bin/bash
# Builds libppp_b200.so (sm_100a) in-tree. Called by __graft_entry__.build(); usable by hand.
set -euo pipefail
cd "$(dirname "$0")/prophylo_b200/csrc"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH"
mkdir -p build
rm -f build/*.o
nvcc $COMMON -c ppp_sweep.cu -o build/ppp_sweep.o &
nvcc $COMMON -fmad=false -c ppp_exact.cu -o build/ppp_exact.o &
nvcc $COMMON -c ppp_topk.cu -o build/ppp_topk.o &
nvcc $COMMON -c ppp_ranked.cu -o build/ppp_ranked.o &
nvcc $COMMON -c ppp_prefilter.cu -o build/ppp_prefilter.o &
nvcc $COMMON -c ppp_prefilter_tc.cu -o build/ppp_prefilter_tc.o &
nvcc $COMMON -fmad=false -c ppp_table.cu -o build/ppp_table.o &
nvcc $COMMON -c ppp_rank.cu -o build/ppp_rank.o &
nvcc $COMMON -c ppp_debug.cu -o build/ppp_debug.o &
nvcc $COMMON -c ppp_capi.cu -o build/ppp_capi.o &
g++ -O2 -std=c++17 -fPIC -pthread -c ppp_bla.cpp -o build/ppp_bla.o &
g++ -O2 -std=c++17 -fPIC -pthread -c ppp_pack.cpp -o build/ppp_pack.o &
FAIL=0; for j in $(jobs -p); do wait $j || FAIL=1; done; [ $FAIL = 0 ] || { echo "nvcc failed"; exit 1; }
nvcc $ARCH -shared -o ../libppp_b200.so build/ppp_sweep.o build/ppp_exact.o build/ppp_topk.o build/ppp_ranked.o build/ppp_prefilter.o build/ppp_prefilter_tc.o build/ppp_table.o build/ppp_rank.o build/ppp_debug.o build/ppp_capi.o build/ppp_bla.o build/ppp_pack.o
echo "built $(cd .. && pwd)/libppp_b200.so"
