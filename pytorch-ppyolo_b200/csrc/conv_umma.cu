// bf16 instantiation of the tcgen05 implicit-GEMM conv (see conv_umma_impl.cuh): ppy_conv_bf16, ppy_conv_bf16_supported.
#define PPY_UMMA_SPLIT 0
#include "conv_umma_impl.cuh"
