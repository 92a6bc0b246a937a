// C-ABI glue: version, error string, convolution engine dispatch.
#include <stdarg.h>
#include "common.cuh"

namespace paif {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int conv_direct_launch(const PaifConvDesc& d, cudaStream_t stream);
int conv_direct_tiles(int H, int W);
int conv_tc_launch(const PaifConvDesc& d, cudaStream_t stream);
int conv_tc_tiles(int H, int W);
bool conv_tc_supported(const PaifConvDesc& d);
int conv_tc_kq(int nsrc, int k, int dil, bool bf16);

}  // namespace paif

using namespace paif;

extern "C" int paif_abi_version(void) { return PAIF_ABI_VERSION; }
extern "C" const char* paif_last_error_string(void) { return g_err; }

extern "C" int paif_conv_forward(const PaifConvDesc* d, void* stream) {
    PAIF_REQUIRE(d != nullptr, "null descriptor");
    PAIF_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0, "bad shape");
    PAIF_REQUIRE(d->nsrc >= 1 && d->nsrc <= 3, "nsrc must be 1..3");
    PAIF_REQUIRE(d->cin_per_src > 0 && d->cin_per_src % 4 == 0, "cin_per_src must be a multiple of 4");
    PAIF_REQUIRE((d->kh & 1) && (d->kw & 1) && d->kh >= 1 && d->kh <= 7 && d->kw >= 1 && d->kw <= 7, "odd kernel <= 7");
    PAIF_REQUIRE(d->dil >= 1 && d->dil <= 2, "dilation 1 or 2");
    PAIF_REQUIRE(d->out != nullptr, "null output");
    for (int i = 0; i < d->nsrc; ++i) PAIF_REQUIRE(d->src[i] != nullptr, "null source");
    PAIF_REQUIRE(d->B <= 65535, "B exceeds grid.z");
    PAIF_REQUIRE(d->storage >= PAIF_STORAGE_F32 && d->storage <= PAIF_STORAGE_F32_BF16, "unknown storage mode");
    int engine = d->engine;
    if (engine == PAIF_ENGINE_AUTO) engine = (d->weight_mma && conv_tc_supported(*d)) ? PAIF_ENGINE_TCGEN05 : PAIF_ENGINE_DIRECT;
    if (engine == PAIF_ENGINE_DIRECT) {
        PAIF_REQUIRE(d->weight != nullptr, "direct engine needs weight");
        PAIF_REQUIRE(d->storage == PAIF_STORAGE_F32, "the direct engine has no bf16 storage mode");
        return conv_direct_launch(*d, (cudaStream_t)stream);
    }
    if (engine == PAIF_ENGINE_TCGEN05) {
        PAIF_REQUIRE(d->weight_mma != nullptr, "tcgen05 engine needs weight_mma");
        if (!conv_tc_supported(*d)) { set_error("paif_conv_forward: shape not supported by the tcgen05 engine"); return PAIF_ENOTSUP; }
        return conv_tc_launch(*d, (cudaStream_t)stream);
    }
    set_error("paif_conv_forward: unknown engine %d", engine);
    return PAIF_EINVAL;
}

extern "C" int paif_conv_num_tiles(int H, int W, int engine) {
    if (engine == PAIF_ENGINE_TCGEN05) return conv_tc_tiles(H, W);
    return conv_direct_tiles(H, W);
}

extern "C" int paif_conv_tc_kq(int nsrc, int k, int dil) {
    if (nsrc < 1 || nsrc > 3 || k < 1 || k > 7 || !(k & 1) || dil < 1 || dil > 2) return 0;
    return conv_tc_kq(nsrc, k, dil, false);
}

extern "C" int paif_conv_tc_kq_bf16(int nsrc, int k, int dil) {
    if (nsrc < 1 || nsrc > 3 || k < 1 || k > 7 || !(k & 1) || dil < 1 || dil > 2) return 0;
    return conv_tc_kq(nsrc, k, dil, true);
}
