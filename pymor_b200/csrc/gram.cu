// Tall-skinny Gram / SYRK  G = A^T diag(w) B  for column panels with n >> k.
//
//   gram_dmma_kernel : TMA (128B-swizzled [128 cols x 16 rows] boxes) -> 6-stage mbarrier ring ->
//                      FP64 tensor-core MMA (mma.sync m8n8k4.f64, SASS DMMA.8x8x4), 128x128 output
//                      tile per CTA held in registers, rows split across CTAs (split-n).
//   gram_simple_kernel: generic DFMA kernel for panels TMA cannot describe (odd / negative /
//                      zero column stride, misaligned base) and for tiny operands.
//   gram_reduce_kernel: fixed-order sum of the split partials, mirror for SYRK, write to `out`
//                      (device or mapped host memory).  Fixed order => bitwise reproducible.
//
// Algorithmic work: 2*n*kA*kB flop (n*k*(k+1) for SYRK), 8*n*(kA+kB) bytes (8*n*k for SYRK).
#include "common.cuh"
#include "gram2_plan.h"
#include "gram_params.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// DMMA kernel
// ------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 6;
constexpr int GRAM_THREADS = 512;                 // 16 warps, 4 (M) x 4 (N), warp tile 32x32
constexpr int TILE_ELEMS = BM * BK;               // 2048 doubles = 16 KB
constexpr uint32_t TILE_BYTES = TILE_ELEMS * 8;
constexpr size_t GRAM_SMEM = 1024 /*align slack*/ + (size_t)STAGES * (2 * TILE_BYTES + BK * 8) +
                             2 * STAGES * sizeof(uint64_t) + 64;

template <bool WEIGHTED>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_dmma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmW, const GramParams p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024B alignment: the 128B swizzle pattern is a function of the absolute smem address
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    double *sA = reinterpret_cast<double *>(smem);
    double *sB = sA + STAGES * TILE_ELEMS;
    double *sW = sB + STAGES * TILE_ELEMS;
    uint64_t *full = reinterpret_cast<uint64_t *>(sW + STAGES * BK);
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    const int item = blockIdx.x;
    const int tile = item % p.ntiles, split = item / p.ntiles;
    int ti, tj;
    decode_tile(p, tile, ti, tj);
    const bool diag = p.symmetric == 1 && ti == tj;   // (symmetric == 2: lower tiles of a general product, two operands)
    const int64_t r0 = (int64_t)split * p.rows_per_split;
    const int64_t r1 = (r0 + p.rows_per_split < p.n) ? r0 + p.rows_per_split : p.n;
    const int nk = (r1 > r0) ? (int)((r1 - r0 + BK - 1) / BK) : 0;

    // Warp -> 32x32 block of the tile.  Each SMSP (warp % 4) has its own FP64 tensor pipe, so when
    // only some blocks carry work they are spread over the SMSPs:
    //  * diagonal SYRK tiles need only the 10 blocks on/below the diagonal: warps 0..9 take them
    //    (3/3/2/2 per SMSP), warps 10..15 only keep the barrier protocol going (measured: -5 %);
    //  * blocks entirely beyond kA / kB are skipped; a tile that is short in M uses the transposed map
    //    so that the surviving blocks sit on all four SMSPs.
    // (Skipping at 8-row granularity inside a block was measured to be SLOWER -- the predicated DMMA
    // sequence issues worse than the straight-line one -- so partially filled blocks run in full.)
    const int vm = min(BM, p.kA - ti * BM), vn = min(BN, p.kB - tj * BN);   // valid rows / cols of this tile
    int wm, wn;
    bool active = true;
    if (diag && !(p.opts & 1)) {
        // (wm, wn) with wn <= wm, enumerated row by row: (0,0) (1,0) (1,1) (2,0) ... (3,3)
        active = warp < 10;
        const int w = active ? warp : 0;
        wm = (w >= 6) ? 3 : (w >= 3) ? 2 : (w >= 1) ? 1 : 0;
        wn = w - wm * (wm + 1) / 2;
    } else if (vm <= BM - 32 && !(p.opts & 4)) {
        wm = warp >> 2; wn = warp & 3;
    } else {
        wm = warp & 3; wn = warp >> 2;
    }
    if (!(p.opts & 2) && (wm * 32 >= vm || wn * 32 >= vn)) active = false;

    if (tid == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (WEIGHTED) tma_prefetch_desc(&tmW);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GRAM_THREADS / 32);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    const uint32_t stage_bytes = TILE_BYTES * (diag ? 1u : 2u) + (WEIGHTED ? BK * 8u : 0u);
    auto issue = [&](int kt) {
        const int s = kt % STAGES;
        const int32_t d = (int32_t)(r0 + (int64_t)kt * BK);
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        tma_load_2d(sA + s * TILE_ELEMS, &tmA, &full[s], d, ti * BM);
        if (!diag) tma_load_2d(sB + s * TILE_ELEMS, &tmB, &full[s], d, tj * BN);
        if (WEIGHTED) tma_load_1d(sW + s * BK, &tmW, &full[s], d);
    };
    if (tid == 0) {
        const int pre = nk < STAGES ? nk : STAGES;
        for (int kt = 0; kt < pre; ++kt) issue(kt);
    }

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    // Fragment row g of an 8-row sub-block is tile row rho(g) = 2*(g&3) + (g>>2).  With the 128B TMA
    // swizzle (16-byte chunk index ^= row & 7) the 16 lanes of each half-warp then hit 16 distinct
    // 8-byte bank pairs: every fragment LDS.64 is conflict-free (2 wavefronts per warp, the minimum).
    const int rho = 2 * (g & 3) + (g >> 2);
    const int h = (lane >> 1) & 1, e = lane & 1;
    int swz[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) swz[ks] = ((((ks * 2 + h) ^ rho) << 1) + e);
    const int rowA = (wm * 32 + rho) * BK;
    const int rowB = (wn * 32 + rho) * BK;

    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % STAGES;
        // producer: refill the stage everybody released one iteration ago
        if (tid == 0 && kt >= 1 && kt - 1 + STAGES < nk) {
            const int sp = (kt - 1) % STAGES;
            mbar_wait(&empty[sp], ((kt - 1) / STAGES) & 1);
            issue(kt - 1 + STAGES);
        }
        mbar_wait(&full[s], (kt / STAGES) & 1);
        if (active) {
            const double *a_s = sA + s * TILE_ELEMS + rowA;
            const double *b_s = (diag ? sA : sB) + s * TILE_ELEMS + rowB;
            const double *w_s = sW + s * BK + t;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = a_s[mi * 8 * BK + swz[ks]];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = b_s[ni * 8 * BK + swz[ks]];
                if (WEIGHTED) {
                    const double wv = w_s[ks * 4];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) b[ni] *= wv;
                }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // epilogue: partial tile, column-major [j][i]; blocks nobody reads later are not written
    if (active) {
        double *out = p.partials + (size_t)item * (BM * BN);
        const int jc0 = 2 * ((2 * t) & 3) + ((2 * t) >> 2);            // rho(2t)
        const int jc1 = 2 * ((2 * t + 1) & 3) + ((2 * t + 1) >> 2);    // rho(2t+1)
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int i = wm * 32 + mi * 8 + rho;
                const int j = wn * 32 + ni * 8;
                out[(size_t)(j + jc0) * BM + i] = acc[mi][ni][0];
                out[(size_t)(j + jc1) * BM + i] = acc[mi][ni][1];
            }
    }
}

// ------------------------------------------------------------------------------------------
// generic kernel: 32x32 output tile, 256 threads (2x2 outputs each), 64 rows per step
// ------------------------------------------------------------------------------------------
constexpr int SM_T = 32, SM_K = 64, SM_PITCH = SM_K + 1, SM_THREADS = 256;

__global__ void __launch_bounds__(SM_THREADS)
gram_simple_kernel(const double *__restrict__ A, int64_t csA, int kA, const double *__restrict__ B,
                   int64_t csB, int kB, const double *__restrict__ w, const GramParams p) {
    __shared__ double As[SM_T * SM_PITCH];
    __shared__ double Bs[SM_T * SM_PITCH];
    const int tid = threadIdx.x;
    const int item = blockIdx.x;
    const int tile = item % p.ntiles, split = item / p.ntiles;
    int ti, tj;
    decode_tile(p, tile, ti, tj);
    const int64_t r0 = (int64_t)split * p.rows_per_split;
    const int64_t r1 = (r0 + p.rows_per_split < p.n) ? r0 + p.rows_per_split : p.n;
    const int tx = tid & 15, ty = tid >> 4;
    double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
    for (int64_t d0 = r0; d0 < r1; d0 += SM_K) {
#pragma unroll
        for (int r = 0; r < (SM_T * SM_K) / SM_THREADS; ++r) {
            const int e = tid + r * SM_THREADS;
            const int ci = e / SM_K, dd = e % SM_K;
            const int64_t d = d0 + dd;
            const int ca = ti * SM_T + ci, cb = tj * SM_T + ci;
            double av = 0.0, bv = 0.0;
            if (d < r1) {
                if (ca < kA) av = A[(int64_t)ca * csA + d];
                if (cb < kB) {
                    bv = B[(int64_t)cb * csB + d];
                    if (w) bv *= w[d];
                }
            }
            As[ci * SM_PITCH + dd] = av;
            Bs[ci * SM_PITCH + dd] = bv;
        }
        __syncthreads();
        const double *a0 = As + (ty * 2) * SM_PITCH, *a1 = a0 + SM_PITCH;
        const double *b0 = Bs + (tx * 2) * SM_PITCH, *b1 = b0 + SM_PITCH;
#pragma unroll 16
        for (int dd = 0; dd < SM_K; ++dd) {
            const double x0 = a0[dd], x1 = a1[dd], y0 = b0[dd], y1 = b1[dd];
            c00 = fma(x0, y0, c00);
            c01 = fma(x0, y1, c01);
            c10 = fma(x1, y0, c10);
            c11 = fma(x1, y1, c11);
        }
        __syncthreads();
    }
    double *out = p.partials + (size_t)item * (SM_T * SM_T);
    const int i = ty * 2, j = tx * 2;
    out[j * SM_T + i] = c00;
    out[(j + 1) * SM_T + i] = c01;
    out[j * SM_T + i + 1] = c10;
    out[(j + 1) * SM_T + i + 1] = c11;
}

// ------------------------------------------------------------------------------------------
// stage 2: out[i + j*ldo] = sum_split partial[split][tile][j][i]   (+ mirror for SYRK)
// ------------------------------------------------------------------------------------------
__global__ void gram_reduce_kernel(const double *__restrict__ partials, int T, int nsplit,
                                   GramParams p, int kA, int kB, double *__restrict__ out,
                                   int64_t ldo, int accumulate) {
    const int tile = blockIdx.y;
    int ti, tj;
    decode_tile(p, tile, ti, tj);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T * T) return;
    const int il = e % T, jl = e / T;
    const int gi = ti * T + il, gj = tj * T + jl;
    if (gi >= kA || gj >= kB) return;
    if (p.symmetric && ti == tj && gi < gj) return;  // strictly-upper part of diagonal tiles: mirrored
    const size_t tile_elems = (size_t)T * T;
    const double *src = partials + (size_t)tile * tile_elems + e;
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) s += src[(size_t)sp * p.ntiles * tile_elems];
    if (accumulate) s += out[gi + (int64_t)gj * ldo];
    out[gi + (int64_t)gj * ldo] = s;
    if (p.symmetric && gi != gj) out[gj + (int64_t)gi * ldo] = s;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int pmc_gram_impl(pmc_ctx *ctx, const double *A, int64_t csA, int32_t kA, const double *B,
                  int64_t csB, int32_t kB, int64_t n, const double *w, double *out, int64_t ldo,
                  uint32_t flags, void *ws, size_t ws_cap, int accumulate) {
    PMC_REQUIRE(ctx && kA >= 0 && kB >= 0 && n >= 0, "bad arguments");
    if (kA == 0 || kB == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(out != nullptr && ldo >= kA, "bad output");
    const bool sym = (flags & PMC_SYMMETRIC) != 0;
    const bool lower = !sym && (flags & PMC_LOWER_ONLY) != 0;
    if (sym) PMC_REQUIRE(A == B && csA == csB && kA == kB, "PMC_SYMMETRIC needs identical panels");
    if (lower) PMC_REQUIRE(kA == kB, "PMC_LOWER_ONLY needs a square result");
    if (n == 0) {
        if (!accumulate)
            PMC_CHECK_CUDA(cudaMemset2DAsync(out, ldo * sizeof(double), 0, kA * sizeof(double), kB, ctx->stream));
        return pmc_after_launch(ctx, 0, flags);
    }
    PMC_REQUIRE(A && B, "NULL panel");

    const bool tma_ok = !(flags & PMC_FORCE_SIMPLE) && csA > 0 && csB > 0 && (csA % 2) == 0 &&
                        (csB % 2) == 0 && aligned16(A) && aligned16(B) && (!w || aligned16(w)) &&
                        n >= BK && n < (int64_t)1 << 31 && csA < ((int64_t)1 << 36) &&
                        csB < ((int64_t)1 << 36) && (kA >= 16 || kB >= 16);
    const int T = tma_ok ? BM : SM_T;
    const int rows_q = tma_ok ? BK : SM_K;

    GramParams p;
    p.n = n;
    p.kA = kA;
    p.kB = kB;
    p.opts = ctx->gram_opts;
    p.symmetric = sym ? 1 : lower ? 2 : 0;
    p.tiles_m = (kA + T - 1) / T;
    p.tiles_n = (kB + T - 1) / T;
    p.ntiles = p.symmetric ? p.tiles_m * (p.tiles_m + 1) / 2 : p.tiles_m * p.tiles_n;

    // split-n: aim at ~gram_waves full waves of CTAs (1 CTA/SM for the DMMA kernel), each item >= 64 steps
    const int64_t steps = ceil_div64(n, rows_q);
    const int ctas_per_wave = ctx->sm_count * (tma_ok ? 1 : 4);
    // items differ in cost (diagonal SYRK tiles are cheaper), so the tail of the last wave matters:
    // more, smaller items keep it short (env PMC_GRAM_WAVES for tuning)
    int64_t nsplit = ceil_div64((int64_t)ctx->gram_waves * ctas_per_wave, p.ntiles);
    // ... and SHORT items keep the operands in L2.  The CTAs that contract the same rows (all tiles of one split)
    // are dispatched one after the other as SMs become free, i.e. staggered by ~0.2 item durations in steady state;
    // the rows the first of them has read must still be in the 126 MB L2 when the last one arrives, for every split
    // in flight (ctas_per_wave / ntiles of them).  Measured at 10M x 1000 (profiles/r02_summary.md): 32 waves of
    // items -> 453 GB read from DRAM for 80 GB of operands, 256 waves -> 188 GB at the same kernel time.
    if (tma_ok && ctx->gram_l2_mb > 0) {
        const double row_bytes = 8.0 * (sym ? kA : kA + kB);                    // distinct operand bytes per row
        const double in_flight = (double)ctas_per_wave / p.ntiles;              // splits being contracted at a time
        double rows = (double)ctx->gram_l2_mb * 1e6 / (0.2 * row_bytes * (in_flight > 1.0 ? in_flight : 1.0));
        if (rows < 64.0 * rows_q) rows = 64.0 * rows_q;
        const int64_t by_l2 = ceil_div64(n, (int64_t)rows);
        // the partial tiles must stay small next to the operands (5 %, and within 4 GB)
        const double ws_limit = 4e9 < 0.05 * row_bytes * n ? 4e9 : 0.05 * row_bytes * n;
        const int64_t by_ws = (int64_t)(ws_limit / ((double)p.ntiles * T * T * 8.0));
        int64_t want = by_l2 < by_ws ? by_l2 : by_ws;
        if (want > nsplit) nsplit = want;
    }
    const int64_t max_split = steps / 64 > 0 ? steps / 64 : 1;
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 1) {
        // round the item count up to a whole number of waves where that is cheap
        int64_t items = nsplit * p.ntiles;
        int64_t waves = ceil_div64(items, ctas_per_wave);
        int64_t ns2 = (waves * ctas_per_wave) / p.ntiles;
        if (ns2 >= 1 && ns2 <= max_split) nsplit = ns2;
    }
    p.rows_per_split = round_up64(ceil_div64(n, nsplit), rows_q);
    nsplit = ceil_div64(n, p.rows_per_split);
    const int64_t items = nsplit * p.ntiles;
    const size_t ws_bytes = (size_t)items * T * T * sizeof(double);
    if (ws_bytes > ws_cap || ws == nullptr) {
        ctx->ws_needed = ws_bytes + ((char *)ws - (char *)ctx->ws);
        pmc_set_error("workspace too small: need %zu bytes, have %zu", ctx->ws_needed, ctx->ws_bytes);
        return PMC_ERR_WORKSPACE;
    }
    int rc;
    p.partials = static_cast<double *>(ws);

    if (tma_ok) {
        CUtensorMap tmA, tmB, tmW;
        rc = pmc_encode_tmap_f64(ctx, &tmA, A, 2, (uint64_t)n, (uint64_t)kA, (uint64_t)csA * 8, BK, BM, true);
        if (rc) return rc;
        rc = pmc_encode_tmap_f64(ctx, &tmB, B, 2, (uint64_t)n, (uint64_t)kB, (uint64_t)csB * 8, BK, BN, true);
        if (rc) return rc;
        if (w) {
            rc = pmc_encode_tmap_f64(ctx, &tmW, w, 1, (uint64_t)n, 1, 0, BK, 1, false);
            if (rc) return rc;
        } else {
            tmW = tmA;
        }
        if (p.opts & G2_OPT_ENABLE) {
            // second-generation tile kernel with its own reduction (gram2.cu)
            rc = pmc_gram2_launch(ctx, tmA, tmB, tmW, p, items, nsplit, w != nullptr, out, ldo, accumulate);
            if (rc) return rc;
            return pmc_after_launch(ctx, 2, flags);
        }
        const int tslot = pmc_timing_begin(ctx, PMC_KERNEL_GRAM);
        if (w) {
            PMC_CHECK_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<true>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAM_SMEM));
            gram_dmma_kernel<true><<<(unsigned)items, GRAM_THREADS, GRAM_SMEM, ctx->stream>>>(tmA, tmB, tmW, p);
        } else {
            PMC_CHECK_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<false>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAM_SMEM));
            gram_dmma_kernel<false><<<(unsigned)items, GRAM_THREADS, GRAM_SMEM, ctx->stream>>>(tmA, tmB, tmW, p);
        }
        pmc_timing_end(ctx, tslot);
    } else {
        gram_simple_kernel<<<(unsigned)items, SM_THREADS, 0, ctx->stream>>>(A, csA, kA, B, csB, kB, w, p);
    }
    PMC_CHECK_CUDA(cudaGetLastError());
    dim3 rgrid((T * T + 255) / 256, p.ntiles);
    gram_reduce_kernel<<<rgrid, 256, 0, ctx->stream>>>(p.partials, T, (int)nsplit, p, kA, kB, out, ldo,
                                                       accumulate);
    return pmc_after_launch(ctx, 2, flags);
}

extern "C" int pmc_gram(pmc_ctx *ctx, const double *A, int64_t csA, int32_t kA, const double *B,
                        int64_t csB, int32_t kB, int64_t n, const double *w, double *out, int64_t ldo,
                        uint32_t flags) {
    PMC_REQUIRE(ctx != nullptr, "ctx is NULL");
    return pmc_gram_impl(ctx, A, csA, kA, B, csB, kB, n, w, out, ldo, flags, ctx->ws, ctx->ws_bytes, 0);
}
