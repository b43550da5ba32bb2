#!/bin/bash
# Builds libmspde.so in-tree for sm_100a (also what __graft_entry__.build() runs).
set -e
cd "$(dirname "$0")/mcmc-multispde_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -diag-suppress 550 \
     -shared -Xcompiler -fPIC -o ../mspde/libmspde.so \
     zgemm_tn.cu potrf.cu geqrf.cu getrf.cu trsm_upper.cu assembly.cu phi.cu pcn.cu tomography.cu probe.cu capi.cu "$@"
