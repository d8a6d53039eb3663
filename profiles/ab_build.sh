#!/bin/bash
# A/B builds of librapid_b200.so that differ in the tuning switches of cd_bucketed.cu (run in the build container):
#   bash profiles/ab_build.sh   ->  rapid_b200/ab/lib_<name>.so ; select with RAPID_B200_LIB=... (same box, same run)
set -e
cd "$(dirname "$0")/.."
mkdir -p rapid_b200/ab
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
OBJS=$(ls rapid_b200/build/*.o | grep -v cd_bucketed.o)
build() {  # name, extra defines
  nvcc $FLAGS $2 -c rapid_b200/csrc/cd_bucketed.cu -o rapid_b200/ab/cd_bucketed_$1.o
  nvcc -shared -o rapid_b200/ab/lib_$1.so $OBJS rapid_b200/ab/cd_bucketed_$1.o -gencode arch=compute_100a,code=sm_100a -ldl
}
build memo0 "-DRAPID_MEMO=0" &
build memo1 "-DRAPID_MEMO=1" &
wait
ls -la rapid_b200/ab/*.so
