// Block lincomb  R = A * C   (n x k) * (k x m) -> (n x m), all panels column-major with the long
// DOF axis contiguous.
//
//   lincomb_dmma_kernel : 128 (rows d) x 128 (cols j) output tile per CTA, reduction over the k
//       snapshots in steps of 16.  A arrives by TMA as eight 128B-swizzled [16 i x 16 d] boxes per
//       stage, the coefficients as one [128 j x 16 i] box (coefficients are stored Fortran-ordered,
//       i.e. K-major, on the device).  The DMMA "A" fragment needs 8 rows (d) x 4 k (i) but d is the
//       contiguous axis, so the 8 rows of a fragment are the permuted set {0,1,8,9,2,3,10,11}(+4):
//       with the 128B swizzle this makes every fragment load bank-conflict free.
//   lincomb_simple_kernel: generic kernel (any stride/alignment; also the right kernel when m is
//       tiny and the op is HBM-bound).
//
// Algorithmic work: 2*n*k*m flop, 8*n*(k+m) bytes.
#include "common.cuh"

namespace {

constexpr int LM = 128, LN = 128, LK = 16;
constexpr int LSTAGES = 6;
constexpr int LTHREADS = 512;
constexpr int A_BOX_ELEMS = LK * 16;               // [16 i][16 d] = 2 KB
constexpr int A_TILE_ELEMS = 8 * A_BOX_ELEMS;      // 16 KB
constexpr int C_TILE_ELEMS = LN * LK;              // 16 KB
constexpr uint32_t L_STAGE_BYTES = (A_TILE_ELEMS + C_TILE_ELEMS) * 8;
constexpr size_t LINCOMB_SMEM = 1024 + (size_t)LSTAGES * L_STAGE_BYTES + 2 * LSTAGES * sizeof(uint64_t) + 64;

struct LincombParams {
    int64_t n;
    int k, m;
    double *R;
    int64_t csR;
    int tiles_j;
};

__global__ void __launch_bounds__(LTHREADS, 1)
lincomb_dmma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmC,
                    const LincombParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    double *sA = reinterpret_cast<double *>(smem);
    double *sC = sA + LSTAGES * A_TILE_ELEMS;
    uint64_t *full = reinterpret_cast<uint64_t *>(sC + LSTAGES * C_TILE_ELEMS);
    uint64_t *empty = full + LSTAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wd = warp & 3, wj = warp >> 2;
    const int g = lane >> 2, t = lane & 3;

    // consecutive CTAs share the same row slab of A (L2 reuse across the m tiles)
    const int64_t tile = blockIdx.x;
    const int tjx = (int)(tile % p.tiles_j);
    const int64_t tdx = tile / p.tiles_j;
    const int64_t d0 = tdx * LM;
    const int j0 = tjx * LN;
    const int nk = (p.k + LK - 1) / LK;

    if (tid == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmC);
        for (int s = 0; s < LSTAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], LTHREADS / 32);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    auto issue = [&](int kt) {
        const int s = kt % LSTAGES;
        const int i0 = kt * LK;
        mbar_arrive_expect_tx(&full[s], L_STAGE_BYTES);
        double *a_dst = sA + s * A_TILE_ELEMS;
#pragma unroll
        for (int b = 0; b < 8; ++b)
            tma_load_2d(a_dst + b * A_BOX_ELEMS, &tmA, &full[s], (int32_t)(d0 + b * 16), i0);
        tma_load_2d(sC + s * C_TILE_ELEMS, &tmC, &full[s], i0, j0);
    };
    if (tid == 0) {
        const int pre = nk < LSTAGES ? nk : LSTAGES;
        for (int kt = 0; kt < pre; ++kt) issue(kt);
    }

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    // A fragment: rows are d = fperm(g) + 4*(mi&1) inside box 2*wd + (mi>>1); k index i = ks*4 + t
    const int fperm = (g & 1) | ((g & 2) << 2) | ((g & 4) >> 1);   // 0,1,8,9,2,3,10,11
    int a_off[4][2];   // [ks][half] offset in doubles inside a [16 i][16 d] box
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = ks * 4 + t;
            const int dd = fperm + 4 * h;
            a_off[ks][h] = i * 16 + ((((dd >> 1) ^ (i & 7)) << 1) | (dd & 1));
        }
    // C fragment: same K-major swizzled pattern as the Gram kernel; fragment column g is tile column
    // rho(g) = 2*(g&3) + (g>>2), which makes each half-warp's 16 loads hit 16 distinct bank pairs
    const int rho = 2 * (g & 3) + (g >> 2);
    int c_off[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) c_off[ks] = ((((ks * 2 + ((lane >> 1) & 1)) ^ rho) << 1) + (lane & 1));
    const int rowC = (wj * 32 + rho) * LK;
    int ni_cnt = (min(LN, p.m - j0) - wj * 32 + 7) >> 3;      // 8-column sub-blocks with data
    ni_cnt = ni_cnt < 0 ? 0 : (ni_cnt > 4 ? 4 : ni_cnt);

    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % LSTAGES;
        if (tid == 0 && kt >= 1 && kt - 1 + LSTAGES < nk) {
            const int sp = (kt - 1) % LSTAGES;
            mbar_wait(&empty[sp], ((kt - 1) / LSTAGES) & 1);
            issue(kt - 1 + LSTAGES);
        }
        mbar_wait(&full[s], (kt / LSTAGES) & 1);
        const double *a_s = sA + s * A_TILE_ELEMS + (2 * wd) * A_BOX_ELEMS;
        const double *c_s = sC + s * C_TILE_ELEMS + rowC;
        if (ni_cnt == 4) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = a_s[(mi >> 1) * A_BOX_ELEMS + a_off[ks][mi & 1]];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = c_s[ni * 8 * LK + c_off[ks]];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
        } else if (ni_cnt > 0) {
            // last column tile when m is not a multiple of 128: skip the 8-column sub-blocks without data
            // (the four short warps sit on four different SMSPs, so the whole CTA gets faster)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = a_s[(mi >> 1) * A_BOX_ELEMS + a_off[ks][mi & 1]];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = c_s[ni * 8 * LK + c_off[ks]];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    if (ni < ni_cnt) {
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
                    }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // epilogue: registers -> global.  One warp store covers 8 full 32-byte sectors.
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int64_t d = d0 + (2 * wd + (mi >> 1)) * 16 + fperm + 4 * (mi & 1);
        if (d >= p.n) continue;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int jb = j0 + wj * 32 + ni * 8;
            const int ja = jb + 2 * ((2 * t) & 3) + ((2 * t) >> 2);          // rho(2t)
            const int jc = jb + 2 * ((2 * t + 1) & 3) + ((2 * t + 1) >> 2);  // rho(2t+1)
            if (ja < p.m) p.R[(int64_t)ja * p.csR + d] = acc[mi][ni][0];
            if (jc < p.m) p.R[(int64_t)jc * p.csR + d] = acc[mi][ni][1];
        }
    }
}

// generic: one thread per row d, JB output columns per thread, coefficients from shared memory
constexpr int SJ = 8, S_THREADS = 256, S_KC = 128;

__global__ void __launch_bounds__(S_THREADS)
lincomb_simple_kernel(const double *__restrict__ A, int64_t csA, int k, int64_t n,
                      const double *__restrict__ CT, int64_t ldc, int m, double *__restrict__ R,
                      int64_t csR) {
    __shared__ double Cs[SJ * S_KC];
    const int j0 = blockIdx.y * SJ;
    const int64_t d = (int64_t)blockIdx.x * S_THREADS + threadIdx.x;
    double acc[SJ];
#pragma unroll
    for (int jj = 0; jj < SJ; ++jj) acc[jj] = 0.0;
    for (int i0 = 0; i0 < k; i0 += S_KC) {
        const int kc = (k - i0 < S_KC) ? k - i0 : S_KC;
        __syncthreads();
        for (int e = threadIdx.x; e < SJ * S_KC; e += S_THREADS) {
            const int jj = e / S_KC, ii = e % S_KC;
            Cs[e] = (j0 + jj < m && ii < kc) ? CT[(int64_t)(j0 + jj) * ldc + i0 + ii] : 0.0;
        }
        __syncthreads();
        if (d < n) {
#pragma unroll 4
            for (int ii = 0; ii < kc; ++ii) {
                const double a = ld_stream1(A + (int64_t)(i0 + ii) * csA + d);
#pragma unroll
                for (int jj = 0; jj < SJ; ++jj) acc[jj] = fma(a, Cs[jj * S_KC + ii], acc[jj]);
            }
        }
    }
    if (d < n) {
#pragma unroll
        for (int jj = 0; jj < SJ; ++jj)
            if (j0 + jj < m) R[(int64_t)(j0 + jj) * csR + d] = acc[jj];
    }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int pmc_lincomb(pmc_ctx *ctx, const double *A, int64_t csA, int32_t k, int64_t n,
                           const double *CT, int64_t ldc, int32_t m, double *R, int64_t csR,
                           uint32_t flags) {
    PMC_REQUIRE(ctx && k >= 0 && m >= 0 && n >= 0, "bad arguments");
    if (m == 0 || n == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(R != nullptr, "NULL output");
    if (k == 0) {
        // empty sum: zero vectors
        PMC_REQUIRE(csR >= n, "output panel must not overlap");
        PMC_CHECK_CUDA(cudaMemset2DAsync(R, csR * sizeof(double), 0, n * sizeof(double), m, ctx->stream));
        return pmc_after_launch(ctx, 0, flags);
    }
    PMC_REQUIRE(A && CT && ldc >= k, "bad operands");

    const bool tma_ok = !(flags & PMC_FORCE_SIMPLE) && csA > 0 && (csA % 2) == 0 && aligned16(A) &&
                        aligned16(CT) && (ldc % 2) == 0 && n >= 16 && n < ((int64_t)1 << 31) &&
                        csA < ((int64_t)1 << 36) && m >= 16 && k >= 4;
    if (tma_ok) {
        CUtensorMap tmA, tmC;
        int rc = pmc_encode_tmap_f64(ctx, &tmA, A, 2, (uint64_t)n, (uint64_t)k, (uint64_t)csA * 8, 16, LK, true);
        if (rc) return rc;
        rc = pmc_encode_tmap_f64(ctx, &tmC, CT, 2, (uint64_t)k, (uint64_t)m, (uint64_t)ldc * 8, LK, LN, true);
        if (rc) return rc;
        LincombParams p;
        p.n = n; p.k = k; p.m = m; p.R = R; p.csR = csR;
        p.tiles_j = (m + LN - 1) / LN;
        const int64_t tiles = ceil_div64(n, LM) * p.tiles_j;
        PMC_REQUIRE(tiles < ((int64_t)1 << 31), "too many tiles");
        PMC_CHECK_CUDA(cudaFuncSetAttribute(lincomb_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)LINCOMB_SMEM));
        const int tslot = pmc_timing_begin(ctx, PMC_KERNEL_LINCOMB);
        lincomb_dmma_kernel<<<(unsigned)tiles, LTHREADS, LINCOMB_SMEM, ctx->stream>>>(tmA, tmC, p);
        pmc_timing_end(ctx, tslot);
    } else {
        dim3 grid((unsigned)ceil_div64(n, S_THREADS), (m + SJ - 1) / SJ);
        lincomb_simple_kernel<<<grid, S_THREADS, 0, ctx->stream>>>(A, csA, k, n, CT, ldc, m, R, csR);
    }
    return pmc_after_launch(ctx, 1, flags);
}
