// Direct fp32 (FFMA) dense convolution engine — the exact-fp32 path of paif_conv_forward and the
// on-device checker for the tcgen05 engine.  Tile = 128 x 8 output pixels per CTA (256 threads);
// warp w owns row w, lane l owns pixels x0 + l + 32*i (i < 4) and all COUT channels, so shared
// memory reads of the input tile are conflict-free and weight reads are warp broadcasts.
// The K loop runs over (source map, channel quad): torch.cat of the reference
// (operations_m.py:446-448) never materialises.
#include "conv_epilogue.cuh"

namespace paif {

constexpr int CD_TW = 128, CD_TH = 8, CD_NT = 256;

struct ConvGeom {
    int B, H, W, nsrc, cin, kh, kw, dil;
    const float* src[3];
    const float* weight;
};

template <int COUT>
__global__ void __launch_bounds__(CD_NT)
conv_direct_kernel(ConvGeom g, EpiParams e) {
    extern __shared__ __align__(16) float smem[];
    const int ph = g.dil * (g.kh - 1) / 2, pw = g.dil * (g.kw - 1) / 2;
    const int RH = CD_TH + 2 * ph, RW = CD_TW + 2 * pw;
    const int taps = g.kh * g.kw;
    float* sIn = smem;                       // [4][RH][RW]
    float* sW = smem + 4 * RH * RW;          // [taps][4][COUT]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = blockIdx.x * CD_TW, y0 = blockIdx.y * CD_TH, b = blockIdx.z;
    const size_t plane = (size_t)g.H * g.W;
    const int Qin = g.cin / 4;

    float acc[4][COUT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[i][c] = 0.f;

    for (int s = 0; s < g.nsrc; ++s) {
        for (int q = 0; q < Qin; ++q) {
            __syncthreads();
            const float4* sp = reinterpret_cast<const float4*>(g.src[s]) + ((size_t)b * Qin + q) * plane;
            for (int idx = tid; idx < RH * RW; idx += CD_NT) {
                const int r = idx / RW, c = idx - r * RW;
                const int yy = y0 - ph + r, xx = x0 - pw + c;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) v = sp[(size_t)yy * g.W + xx];
                sIn[0 * RH * RW + idx] = v.x;
                sIn[1 * RH * RW + idx] = v.y;
                sIn[2 * RH * RW + idx] = v.z;
                sIn[3 * RH * RW + idx] = v.w;
            }
            // weights [s][tap][cin][COUT] -> sW[tap][4][COUT]
            for (int idx = tid; idx < taps * COUT; idx += CD_NT) {   // float4 granules
                const int t = idx / COUT, rem = idx - t * COUT;      // rem: 4*COUT/4 float4 per tap
                const float4 v = reinterpret_cast<const float4*>(
                    g.weight + (((size_t)s * taps + t) * g.cin + q * 4) * COUT)[rem];
                reinterpret_cast<float4*>(sW + (size_t)t * 4 * COUT)[rem] = v;
            }
            __syncthreads();
            for (int ty = 0; ty < g.kh; ++ty) {
                for (int tx = 0; tx < g.kw; ++tx) {
                    const float* wt = sW + (size_t)(ty * g.kw + tx) * 4 * COUT;
                    const float* in0 = sIn + (warp + ty * g.dil) * RW + lane + tx * g.dil;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float a[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) a[i] = in0[c * RH * RW + 32 * i];
#pragma unroll
                        for (int o4 = 0; o4 < COUT / 4; ++o4) {
                            const float4 wv = reinterpret_cast<const float4*>(wt + c * COUT)[o4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                acc[i][o4 * 4 + 0] = fmaf(a[i], wv.x, acc[i][o4 * 4 + 0]);
                                acc[i][o4 * 4 + 1] = fmaf(a[i], wv.y, acc[i][o4 * 4 + 1]);
                                acc[i][o4 * 4 + 2] = fmaf(a[i], wv.z, acc[i][o4 * 4 + 2]);
                                acc[i][o4 * 4 + 3] = fmaf(a[i], wv.w, acc[i][o4 * 4 + 3]);
                            }
                        }
                    }
                }
            }
        }
    }

    // epilogue
    const int y = y0 + warp;
    float csum[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) csum[c] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x0 + lane + 32 * i;
        if (y < g.H && x < g.W) {
            epilogue_pixel<COUT>(e, b, y, x, acc[i]);
#pragma unroll
            for (int c = 0; c < COUT; ++c) csum[c] += acc[i][c];
        }
    }
    if (e.chan_partials) {
        __syncthreads();
        float* red = smem;   // [8 warps][COUT]
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
            float v = csum[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp * COUT + c] = v;
        }
        __syncthreads();
        if (tid < COUT) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < CD_TH; ++w) v += red[w * COUT + tid];
            const int tiles = gridDim.x * gridDim.y;
            const int tile = blockIdx.y * gridDim.x + blockIdx.x;
            e.chan_partials[((size_t)b * tiles + tile) * COUT + tid] = v;
        }
    }
}

int conv_direct_launch(const PaifConvDesc& d, cudaStream_t stream) {
    ConvGeom g;
    g.B = d.B; g.H = d.H; g.W = d.W; g.nsrc = d.nsrc; g.cin = d.cin_per_src;
    g.kh = d.kh; g.kw = d.kw; g.dil = d.dil;
    for (int i = 0; i < 3; ++i) g.src[i] = static_cast<const float*>(d.src[i]);
    g.weight = d.weight;
    EpiParams e = make_epi(d);
    const int ph = d.dil * (d.kh - 1) / 2, pw = d.dil * (d.kw - 1) / 2;
    const int RH = CD_TH + 2 * ph, RW = CD_TW + 2 * pw;
    size_t smem = (size_t)(4 * RH * RW + d.kh * d.kw * 4 * d.cout) * sizeof(float);
    if (smem < (size_t)CD_TH * d.cout * sizeof(float)) smem = (size_t)CD_TH * d.cout * sizeof(float);
    dim3 grid(cdiv(d.W, CD_TW), cdiv(d.H, CD_TH), d.B);
    cudaError_t err;
    if (d.cout == 32) {
        err = cudaFuncSetAttribute(conv_direct_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) { set_error("conv_direct smem attr: %s", cudaGetErrorString(err)); return (int)err; }
        conv_direct_kernel<32><<<grid, CD_NT, smem, stream>>>(g, e);
    } else if (d.cout == 16) {
        err = cudaFuncSetAttribute(conv_direct_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) { set_error("conv_direct smem attr: %s", cudaGetErrorString(err)); return (int)err; }
        conv_direct_kernel<16><<<grid, CD_NT, smem, stream>>>(g, e);
    } else {
        set_error("conv_direct: cout must be 16 or 32");
        return PAIF_ENOTSUP;
    }
    return check_launch("paif_conv_forward(direct)");
}

int conv_direct_tiles(int H, int W) { return cdiv(W, CD_TW) * cdiv(H, CD_TH); }

}  // namespace paif
