#!/bin/bash
# verbose build of the library (registers / spills per kernel go to /tmp/ptxas.log)
cd /root/repo/multigrid_dolfinx_b200 && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fopenmp,-O3 -shared -Xptxas -v -o libmgb200.so csrc/mgb_engine.cu csrc/mgb_setup.cpp 2> /tmp/ptxas.log
echo "exit $?"
grep -n "error" /tmp/ptxas.log | head
