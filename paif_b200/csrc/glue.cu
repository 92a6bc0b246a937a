// Kernels either side of the fusion hot path: the PGD step of attack/attack.py:504-512.
#include "common.cuh"

namespace paif {

// delta <- clamp(clamp(delta + alpha * sign(grad), -eps, eps), 0 - x, 1 - x), in place, for one image tensor.
// attack/attack.py:504-512 does this with six elementwise launches per modality (sign, mul-add, three clamps, the
// .data assignment); `grad` is delta.grad, which the reference never zeroes (it steps along the sign of the RUNNING
// SUM of gradients) — the accumulation stays with autograd, this kernel only reads it.
// torch.sign(0) == 0 and torch.clamp(x, min=lo, max=hi) == min(max(x, lo), hi) are followed exactly.
__global__ void __launch_bounds__(256)
pgd_step_kernel(float* __restrict__ delta, const float* __restrict__ grad, const float* __restrict__ x,
                float alpha, float eps, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n) {
            float4 d = ld4(delta + i);
            const float4 g = ld4(grad + i), xv = ld4(x + i);
            float* dp = &d.x;
            const float* gp = &g.x;
            const float* xp = &xv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float s = gp[k] > 0.f ? 1.f : (gp[k] < 0.f ? -1.f : 0.f);
                float v = __fadd_rn(dp[k], __fmul_rn(alpha, s));
                v = fminf(fmaxf(v, -eps), eps);
                dp[k] = fminf(fmaxf(v, __fsub_rn(0.f, xp[k])), __fsub_rn(1.f, xp[k]));
            }
            st4(delta + i, d);
        } else {
            for (long long j = i; j < n; ++j) {
                const float gj = grad[j];
                const float s = gj > 0.f ? 1.f : (gj < 0.f ? -1.f : 0.f);
                float v = __fadd_rn(delta[j], __fmul_rn(alpha, s));
                v = fminf(fmaxf(v, -eps), eps);
                delta[j] = fminf(fmaxf(v, __fsub_rn(0.f, x[j])), __fsub_rn(1.f, x[j]));
            }
        }
    }
}

}  // namespace paif

using namespace paif;

extern "C" int paif_pgd_step(float* delta, const float* grad, const float* x, float alpha, float eps,
                             long long n, void* stream) {
    PAIF_REQUIRE(delta && grad && x, "null pointer");
    PAIF_REQUIRE(n >= 0, "negative size");
    PAIF_REQUIRE(((reinterpret_cast<uintptr_t>(delta) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(x)) & 15) == 0,
                 "pointers must be 16-byte aligned");
    if (n == 0) return 0;
    long long want = (n / 4 + 255) / 256;
    const int blocks = (int)(want < 148 * 8 ? (want < 1 ? 1 : want) : 148 * 8);
    pgd_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(delta, grad, x, alpha, eps, n);
    return check_launch("paif_pgd_step");
}
