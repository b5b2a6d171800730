#!/bin/sh
# profiling build of the library (phase cycle counters of the analysis stage): _exp/libpsi_prof.so
set -e
cd "$(dirname "$0")/../psitools_public_b200/csrc"
mkdir -p ../../_exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC,-fopenmp -lgomp -ldl \
     -DPSI_PROF -o ../../_exp/libpsi_prof.so psi_b200.cu psi_engine.cu psi_refine.cu
