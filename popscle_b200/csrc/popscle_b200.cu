// popscle_b200.cu — single translation unit of libpopscle_b200.so (C ABI: include/popscle_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include "common.cuh"

#include "context.inl"
#include "demux.inl"
#include "freemux.inl"
#include "multi.inl"
