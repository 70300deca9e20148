// fp32-grade instantiation of the tcgen05 implicit-GEMM conv (see conv_umma_impl.cuh): activations and weights are fp16
// hi/lo pairs (two planes), every K block runs three MMAs (hi*hi + hi*lo + lo*hi) into the same fp32 TMEM accumulator:
// ppy_conv_f16x2.
#define PPY_UMMA_SPLIT 1
#include "conv_umma_impl.cuh"
