// Roofline denominators measured by the library itself, in the process and under the clocks of the run that quotes them:
// the FP64 DMMA (mma.sync.m8n8k4) and DFMA issue peaks of the context's device.  `tcgen05` has no FP64 and the larger
// mma.sync .f64 shapes (m16n8k4/k8/k16) compile to sequences of DMMA.8x8x4 on sm_100a (cuobjdump -sass), so m8n8k4 is the
// hardware primitive whose issue rate bounds every dense contraction step of the hot path.
#include <algorithm>

#include "tne_internal.h"

namespace {

template <int ILP>
__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
    const double av = a + threadIdx.x * 1e-6, bv = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(av), "d"(bv));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

}  // namespace

extern "C" int tne_measure_fp64_peak(tne_ctx* ctx, double* dmma_tflops, double* dfma_tflops) {
    TNE_ENTER(ctx);
    if (ctx->plan_only) return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "planner context: nothing can be measured");
    tne_buf h;
    TNE_TRY(buf_new(ctx, {64}, &h));
    double* d = reinterpret_cast<double*>(buf_get(ctx, h)->ptr());
    cudaEvent_t e0, e1;
    TNE_CUDA(ctx, cudaEventCreate(&e0));
    TNE_CUDA(ctx, cudaEventCreate(&e1));
    const int blocks = ctx->sm_count * 8;
    constexpr int ILP = 8;
    double best[2] = {0.0, 0.0};
    for (int which = 0; which < 2; ++which) {
        const int iters = which == 0 ? (1 << 14) : (1 << 16);
        for (int r = 0; r < 6; ++r) {   // the first repetition is the warm-up
            TNE_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
            if (which == 0) peak_dmma_kernel<ILP><<<blocks, 256, 0, ctx->stream>>>(d, iters, 1.0000001, 1e-9);
            else peak_dfma_kernel<ILP><<<blocks, 256, 0, ctx->stream>>>(d, iters, 1.0000001, 1e-9);
            TNE_LAUNCH_CHECK(ctx);
            TNE_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
            TNE_CUDA(ctx, cudaEventSynchronize(e1));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fl = which == 0 ? 2.0 * 8 * 8 * 4 * (double)ILP * iters * (256.0 / 32) * blocks : 2.0 * ILP * (double)iters * 256.0 * blocks;
            if (r > 0 && ms > 0.f) best[which] = std::max(best[which], fl / (ms * 1e-3) / 1e12);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    tne_free(ctx, h);
    if (dmma_tflops) *dmma_tflops = best[0];
    if (dfma_tflops) *dfma_tflops = best[1];
    return TNE_OK;
}
