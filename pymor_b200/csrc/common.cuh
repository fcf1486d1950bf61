// Shared host/device helpers for libpymor_b200 (sm_100a only).
#pragma once
#include <cuda.h>          // CUtensorMap types only; the driver is reached via cudaGetDriverEntryPoint
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pymor_b200.h"

// ----------------------------------------------------------------------------------------------
// context
// ----------------------------------------------------------------------------------------------
#define PMC_TIMING_SLOTS 1024

struct pmc_ctx {
    int device;
    cudaStream_t stream;
    int sm_count;
    void *ws;             // caller-owned device scratch
    size_t ws_bytes;
    size_t ws_needed;
    int64_t launches;
    // small library-owned buffers (not vector data): ticket counters for single-launch reductions
    unsigned int *tickets;   // device, zero-initialised, self-resetting
    int n_tickets;
    double *partials;        // device, block partials of single-launch reductions
    int partials_cap;        // in doubles; followed by as many int64 (amax indices)
    void *encode_tiled;      // cuTensorMapEncodeTiled entry point
    int csr_variant;         // 0 = row-per-thread SpMM (default; 3.4 TB/s on the 5-point stencil), 1 = row-tile variant (slower, kept for A/B)
    int gram_waves;          // target number of CTA waves of the split-n Gram kernel (env PMC_GRAM_WAVES)
    int gram_l2_mb;          // L2 budget (MB) for the rows shared by the tiles of the splits in flight; 0: split by waves only
    int gram_opts;           // ablation switches for gram_dmma_kernel (env PMC_GRAM_OPTS)
    int l2_keep;             // keep single hot vectors L2-resident (env PMC_NO_L2_KEEP=1 disables)
    // Gram-Schmidt sweep kernel (mgs.cu): flagged CTA packets, abort flag, step counter; peer inboxes
    uint64_t *sweep_box;
    unsigned int *sweep_abort;
    uint32_t sweep_seq;
    int peer_rank, peer_size, peer_pending_size;
    uint64_t *peer_inbox_local;
    uint64_t *peer_inbox[PMC_MAX_PEERS];
    // NCCL communicator of the row-sharded arrays (comm.cu); NULL: single GPU or collectives done by the host
    void *nccl_comm;
    int comm_rank, comm_size;
    // per-launch kernel times (option "kernel_timing"): event pairs around the launches of the hot kernels
    int kernel_timing;
    int t_count;
    int t_kind[PMC_TIMING_SLOTS];
    cudaEvent_t t_ev[PMC_TIMING_SLOTS][2];
};

// Bracket ONE kernel launch with events on the launching stream when the option "kernel_timing" is on
// (bench.py's roofline: the time of the kernel itself, not of the call around it).  begin returns the slot
// (-1: timing off or log full); end ignores -1.
int pmc_timing_begin(pmc_ctx *ctx, int kind);
void pmc_timing_end(pmc_ctx *ctx, int slot);

void pmc_set_error(const char *fmt, ...);

#define PMC_CHECK_CUDA(expr)                                                                     \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            pmc_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return PMC_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

#define PMC_REQUIRE(cond, msg)                                         \
    do {                                                               \
        if (!(cond)) {                                                 \
            pmc_set_error("%s:%d: %s", __FILE__, __LINE__, msg);       \
            return PMC_ERR_INVALID;                                    \
        }                                                              \
    } while (0)

static inline int pmc_after_launch(pmc_ctx *ctx, int nlaunch, uint32_t flags) {
    ctx->launches += nlaunch;
    PMC_CHECK_CUDA(cudaGetLastError());
    if (flags & PMC_SYNC) PMC_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PMC_OK;
}

static inline int pmc_need_ws(pmc_ctx *ctx, size_t bytes) {
    if (bytes > ctx->ws_bytes || (bytes > 0 && ctx->ws == nullptr)) {
        ctx->ws_needed = bytes;
        pmc_set_error("workspace too small: need %zu bytes, have %zu", bytes, ctx->ws_bytes);
        return PMC_ERR_WORKSPACE;
    }
    return PMC_OK;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up64(int64_t a, int64_t b) { return ceil_div64(a, b) * b; }

// Encode a tiled fp64 tensor map (rank 1 or 2).  dims/box in elements, stride1 in bytes.
int pmc_encode_tmap_f64(pmc_ctx *ctx, CUtensorMap *tm, const void *base, int rank, uint64_t dim0,
                        uint64_t dim1, uint64_t stride1_bytes, uint32_t box0, uint32_t box1,
                        bool swizzle128);

// pmc_gram with an explicit scratch region and optional accumulation into `out` (csr.cu uses it
// panel by panel).
int pmc_gram_impl(pmc_ctx *ctx, const double *A, int64_t csA, int32_t kA, const double *B, int64_t csB,
                  int32_t kB, int64_t n, const double *w, double *out, int64_t ldo, uint32_t flags,
                  void *ws, size_t ws_cap, int accumulate);

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// TMA tiled loads global -> shared, completion signalled on an mbarrier (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, uint64_t *bar,
                                            int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const CUtensorMap *tm, uint64_t *bar,
                                            int32_t c0) {
    asm volatile(
        "cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// FP64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col)   (SASS: DMMA.8x8x4)
// fragment ownership (lane = 4*g + t):  a = A[g][t],  b = B[t][g],  c0/c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// streaming 128-bit loads/stores (vector data is touched once per kernel: keep it out of L1)
__device__ __forceinline__ double2 ld_stream2(const double *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream1(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// L2 eviction-priority policies (126 MB L2): a vector that is re-used by the next kernels (the
// column being orthogonalised in Gram-Schmidt) is loaded/stored evict_last, the operands streaming
// past it evict_first, so that it stays L2-resident across launches.
__device__ __forceinline__ uint64_t l2_policy(int kind /*0 normal, 1 evict_first, 2 evict_last*/) {
    uint64_t pol;
    if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double2 ld_hint2_nc(const double *p, uint64_t pol) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double2 ld_hint2(const double *p, uint64_t pol) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p), "l"(pol)
                 : "memory");
    return v;
}
__device__ __forceinline__ void st_hint2(double *p, double2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol)
                 : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__
