// Roofline denominators: sustained FP64 throughput of the tensor-core path (DMMA.8x8x4) and of the
// plain DFMA pipe, measured with register-resident operands on every SM.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(512, 1) dmma_peak_kernel(int iters, double *sink) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(512, 1) dfma_peak_kernel(int iters, double *sink) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    const double a = 1.0 + 1e-12 * threadIdx.x, b = 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) sink[0] = s;
}

// Are the DMMA and the DFMA datapath the same FP64 units?  Mixed issue: mode 2 = half of the warps of every SMSP
// run the DMMA loop, the other half the DFMA loop; mode 3 = every warp interleaves 16 DMMA with 16 DFMA.
// If the combined rate exceeds the larger single rate, a hybrid tile kernel could use both.
__global__ void __launch_bounds__(512, 1) mixed_peak_kernel(int iters, int interleave, double *sink) {
    double c[16][2], d[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; d[i] = threadIdx.x * 1e-3 + i; }
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9, e = 1e-12;
    const bool do_mma = interleave || (((threadIdx.x >> 5) >> 2) & 1) == 0;   // warps 0-3, 8-11: one pair per SMSP
    const bool do_fma = interleave || !do_mma;
    for (int it = 0; it < iters; ++it) {
        if (do_mma) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
        if (do_fma) {
#pragma unroll
            for (int i = 0; i < 16; ++i) d[i] = fma(d[i], a, e);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1] + d[i];
    if (s == 123.456) sink[0] = s;
}

}  // namespace

extern "C" int pmc_measure_fp64_peak(pmc_ctx *ctx, int use_dmma, double ms, double *tflops_out) {
    PMC_REQUIRE(ctx && tflops_out && ms > 0, "bad arguments");
    cudaEvent_t e0, e1;
    PMC_CHECK_CUDA(cudaEventCreate(&e0));
    PMC_CHECK_CUDA(cudaEventCreate(&e1));
    const int grid = ctx->sm_count, threads = 512;
    int iters = 2000;
    double best = 0.0;
    float t = 0.f;
    // calibrate so that one launch runs for about `ms`, then take the best of 3 launches
    for (int round = 0; round < 5; ++round) {
        PMC_CHECK_CUDA(cudaEventRecord(e0, ctx->stream));
        if (use_dmma >= 2) mixed_peak_kernel<<<grid, threads, 0, ctx->stream>>>(iters, use_dmma == 3, ctx->partials);
        else if (use_dmma) dmma_peak_kernel<<<grid, threads, 0, ctx->stream>>>(iters, ctx->partials);
        else dfma_peak_kernel<<<grid, threads, 0, ctx->stream>>>(iters, ctx->partials);
        PMC_CHECK_CUDA(cudaEventRecord(e1, ctx->stream));
        PMC_CHECK_CUDA(cudaEventSynchronize(e1));
        PMC_CHECK_CUDA(cudaEventElapsedTime(&t, e0, e1));
        ctx->launches += 1;
        const double mma_flop = 16.0 * 512.0 /*8x8x4 FMA*/ * (threads / 32), fma_flop = 16.0 * 2.0 * threads;
        const double flop_per_iter = use_dmma == 3 ? mma_flop + fma_flop           // every warp does both
                                     : use_dmma == 2 ? 0.5 * (mma_flop + fma_flop)  // half the warps each
                                     : use_dmma ? mma_flop : fma_flop;
        const double tf = flop_per_iter * iters * grid / (t * 1e-3) * 1e-12;
        if (round >= 2 && tf > best) best = tf;
        if (round < 2) {
            double scale = ms / (t > 1e-3 ? t : 1e-3);
            if (scale > 64) scale = 64;
            iters = (int)(iters * scale) + 1;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops_out = best;
    return PMC_OK;
}
