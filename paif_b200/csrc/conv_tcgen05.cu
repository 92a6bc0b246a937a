// tcgen05 TF32 implicit-GEMM convolution engine (placeholder until the kernel lands).
#include "conv_epilogue.cuh"
namespace paif {
bool conv_tc_supported(const PaifConvDesc&) { return false; }
int conv_tc_tiles(int, int) { return 0; }
int conv_tc_launch(const PaifConvDesc&, cudaStream_t) { set_error("tcgen05 engine not built"); return PAIF_ENOTSUP; }
}
