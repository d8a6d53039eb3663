#!/bin/bash
# A/B builds of librapid_b200.so that differ in the tuning switches of cd_bucketed.cu (run in the build container):
#   bash profiles/ab_build.sh   ->  rapid_b200/ab/lib_<name>.so ; select with RAPID_B200_LIB=... (same box, same run)
set -e
cd "$(dirname "$0")/.."
mkdir -p rapid_b200/ab
rm -f rapid_b200/ab/*
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
OBJS=$(ls rapid_b200/build/*.o | grep -v cd_bucketed.o)
build() {  # name, extra defines
  nvcc $FLAGS $2 -Xptxas -v -c rapid_b200/csrc/cd_bucketed.cu -o rapid_b200/ab/cd_bucketed_$1.o 2> rapid_b200/ab/ptxas_$1.txt
  nvcc -shared -o rapid_b200/ab/lib_$1.so $OBJS rapid_b200/ab/cd_bucketed_$1.o -gencode arch=compute_100a,code=sm_100a -ldl
}
build pf0 "-DRAPID_PF=0" &
build pf2 "-DRAPID_PF=2" &
build pf4 "-DRAPID_PF=4" &
build pf8 "-DRAPID_PF=8" &
build pf4l1 "-DRAPID_PF=4 -DRAPID_PF_L1=1" &
wait
ls -la rapid_b200/ab/*.so
