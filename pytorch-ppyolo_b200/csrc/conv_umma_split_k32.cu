// fp32-grade (fp16 hi/lo pair) instantiation of the tcgen05 conv with 32-element K blocks (64-byte operand rows, SWIZZLE_64B):
// ppy_conv_f16x2_k32, the 3x3 convs with 32 input channels (stem conv1_2 / conv1_3).  See conv_umma_impl.cuh.
#define PPY_UMMA_SPLIT 1
#define PPY_UMMA_K32 1
#include "conv_umma_impl.cuh"
