// CSR operator kernels: Y = M X on column panels, and the fused two-form V^T (M U) that never
// materialises the n x k product (row panels of M U live in the workspace and are contracted by
// the Gram kernel while still L2-resident).
//
// Algorithmic work: apply: 2*nnz*k flop, 12*nnz + 16*n*k bytes;  two-form: + 2*n*kV*kU flop.
#include "common.cuh"

namespace {

constexpr int CJ = 8;          // panel columns per thread (matrix entries are read once per CJ columns)
constexpr int CSR_THREADS = 256;

__global__ void __launch_bounds__(CSR_THREADS)
csr_apply_kernel(int64_t row0, int64_t nrows, const int64_t *__restrict__ rowptr,
                 const int32_t *__restrict__ colidx, const double *__restrict__ vals,
                 const double *__restrict__ X, int64_t csX, int64_t n_local,
                 const double *__restrict__ ghost, double *__restrict__ Y, int64_t csY, int k) {
    const int64_t r = (int64_t)blockIdx.x * CSR_THREADS + threadIdx.x;
    if (r >= nrows) return;
    const int j0 = blockIdx.y * CJ;
    const int64_t p0 = rowptr[row0 + r], p1 = rowptr[row0 + r + 1];
    double acc[CJ];
#pragma unroll
    for (int jj = 0; jj < CJ; ++jj) acc[jj] = 0.0;
    // column indices >= n_local address the halo: rows owned by other ranks, received into
    // ghost[(col - n_local) * k + j] (row-major, one row per ghost DOF)
    for (int64_t p = p0; p < p1; ++p) {
        const double v = vals[p];
        const int64_t col = colidx[p];
        if (col < n_local) {
            const double *x = X + col + (int64_t)j0 * csX;
            if (j0 + CJ <= k) {
#pragma unroll
                for (int jj = 0; jj < CJ; ++jj) acc[jj] = fma(v, x[(int64_t)jj * csX], acc[jj]);
            } else {
#pragma unroll
                for (int jj = 0; jj < CJ; ++jj)
                    if (j0 + jj < k) acc[jj] = fma(v, x[(int64_t)jj * csX], acc[jj]);
            }
        } else {
            const double *gx = ghost + (col - n_local) * k + j0;
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj)
                if (j0 + jj < k) acc[jj] = fma(v, gx[jj], acc[jj]);
        }
    }
#pragma unroll
    for (int jj = 0; jj < CJ; ++jj)
        if (j0 + jj < k) Y[r + (int64_t)(j0 + jj) * csY] = acc[jj];
}

// Row-tile variant: the block's slice of (colidx, vals) is staged in shared memory with coalesced
// loads, so the gathers of X are not queued behind dependent matrix loads; every thread owns one row
// and walks CT_COLS panel columns four at a time (four independent gathers per matrix entry).
constexpr int CT_ROWS = 128, CT_COLS = 32, CT_MAXNZ = CT_ROWS * 24;

__global__ void __launch_bounds__(CT_ROWS)
csr_apply_tiled_kernel(int64_t row0, int64_t nrows, const int64_t *__restrict__ rowptr,
                       const int32_t *__restrict__ colidx, const double *__restrict__ vals,
                       const double *__restrict__ X, int64_t csX, int64_t n_local,
                       const double *__restrict__ ghost, double *__restrict__ Y, int64_t csY, int k) {
    __shared__ int32_t s_col[CT_MAXNZ];
    __shared__ double s_val[CT_MAXNZ];
    const int64_t rb = (int64_t)blockIdx.x * CT_ROWS;
    const int rows = (int)((nrows - rb < CT_ROWS) ? nrows - rb : CT_ROWS);
    const int64_t pb = rowptr[row0 + rb], pe = rowptr[row0 + rb + rows];
    const int cnt = (int)(pe - pb);
    const bool staged = cnt <= CT_MAXNZ;
    if (staged) {
        for (int i = threadIdx.x; i < cnt; i += CT_ROWS) {
            s_col[i] = colidx[pb + i];
            s_val[i] = vals[pb + i];
        }
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= rows) return;
    const int64_t r = rb + t;
    const int64_t p0 = rowptr[row0 + r], p1 = rowptr[row0 + r + 1];
    const int j_begin = blockIdx.y * CT_COLS;
    const int j_end = (j_begin + CT_COLS < k) ? j_begin + CT_COLS : k;
    for (int j = j_begin; j < j_end; j += 4) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const int nj = j_end - j;
        for (int64_t p = p0; p < p1; ++p) {
            const double v = staged ? s_val[p - pb] : vals[p];
            const int64_t col = staged ? s_col[p - pb] : colidx[p];
            if (col < n_local) {
                const double *x = X + col + (int64_t)j * csX;
                a0 = fma(v, x[0], a0);
                if (nj > 1) a1 = fma(v, x[csX], a1);
                if (nj > 2) a2 = fma(v, x[2 * csX], a2);
                if (nj > 3) a3 = fma(v, x[3 * csX], a3);
            } else {
                const double *gx = ghost + (col - n_local) * k + j;
                a0 = fma(v, gx[0], a0);
                if (nj > 1) a1 = fma(v, gx[1], a1);
                if (nj > 2) a2 = fma(v, gx[2], a2);
                if (nj > 3) a3 = fma(v, gx[3], a3);
            }
        }
        double *y = Y + r + (int64_t)j * csY;
        y[0] = a0;
        if (nj > 1) y[csY] = a1;
        if (nj > 2) y[2 * csY] = a2;
        if (nj > 3) y[3 * csY] = a3;
    }
}

// halo packing: out[r * k + j] = A[idx[r] + j * csA]   (idx on the device)
__global__ void gather_rows_kernel(const double *__restrict__ A, int64_t csA, int k,
                                   const int64_t *__restrict__ idx, int64_t nidx, double *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nidx * k) return;
    const int64_t r = t / k;
    const int j = (int)(t % k);
    out[t] = A[idx[r] + (int64_t)j * csA];
}

// Fused pairwise two-form  out[j] = sum_r V[r, j] * (M U)[r, j]  (Operator.pairwise_apply2 with a CSR operator, e.g.
// the product norms ||u_j||_M^2 of every Gram-Schmidt step): the row of M U is formed in registers, multiplied with
// the matching entry of V and reduced -- the n x k product M U is never written.  Every block owns a FIXED set of
// row chunks and writes one partial per column; the second kernel sums the block partials in a fixed tree, so the
// result is bitwise reproducible.
constexpr int CP_MAXBLOCKS = 8 * 148;

__global__ void __launch_bounds__(CSR_THREADS)
csr_pairwise_kernel(int64_t nrows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                    const double *__restrict__ vals, const double *__restrict__ V, int64_t csV,
                    const double *__restrict__ U, int64_t csU, const double *__restrict__ ghost, int k,
                    double *__restrict__ partials /* [gridDim.y][gridDim.x][CJ] */) {
    __shared__ double red[CSR_THREADS / 32][CJ];
    const int j0 = blockIdx.y * CJ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double tot[CJ];
#pragma unroll
    for (int jj = 0; jj < CJ; ++jj) tot[jj] = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * CSR_THREADS + threadIdx.x; r < nrows;
         r += (int64_t)gridDim.x * CSR_THREADS) {
        const int64_t p0 = rowptr[r], p1 = rowptr[r + 1];
        double acc[CJ];
#pragma unroll
        for (int jj = 0; jj < CJ; ++jj) acc[jj] = 0.0;
        for (int64_t p = p0; p < p1; ++p) {
            const double v = vals[p];
            const int64_t col = colidx[p];
            if (col < nrows) {
                const double *x = U + col + (int64_t)j0 * csU;
#pragma unroll
                for (int jj = 0; jj < CJ; ++jj)
                    if (j0 + jj < k) acc[jj] = fma(v, x[(int64_t)jj * csU], acc[jj]);
            } else {
                const double *gx = ghost + (col - nrows) * k + j0;
#pragma unroll
                for (int jj = 0; jj < CJ; ++jj)
                    if (j0 + jj < k) acc[jj] = fma(v, gx[jj], acc[jj]);
            }
        }
#pragma unroll
        for (int jj = 0; jj < CJ; ++jj)
            if (j0 + jj < k) tot[jj] = fma(ld_stream1(V + r + (int64_t)(j0 + jj) * csV), acc[jj], tot[jj]);
    }
#pragma unroll
    for (int jj = 0; jj < CJ; ++jj) {
        const double s = warp_sum(tot[jj]);
        if (lane == 0) red[warp][jj] = s;
    }
    __syncthreads();
    if (threadIdx.x < CJ) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < CSR_THREADS / 32; ++w) s += red[w][threadIdx.x];
        partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * CJ + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256)
csr_pairwise_reduce_kernel(const double *__restrict__ partials, int nblocks, int k, double *__restrict__ out) {
    __shared__ double red[256];
    const int j = blockIdx.x;
    const double *src = partials + ((size_t)(j / CJ) * nblocks) * CJ + (j % CJ);
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) s += src[(size_t)b * CJ];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[j] = red[0];
}

int launch_csr_apply(pmc_ctx *ctx, int64_t row0, int64_t nrows, const int64_t *rowptr,
                     const int32_t *colidx, const double *vals, const double *X, int64_t csX,
                     int64_t n_local, const double *ghost, double *Y, int64_t csY, int k) {
    const int tslot = pmc_timing_begin(ctx, PMC_KERNEL_SPMM);
    if (ctx->csr_variant == 1) {
        dim3 grid((unsigned)ceil_div64(nrows, CT_ROWS), (k + CT_COLS - 1) / CT_COLS);
        csr_apply_tiled_kernel<<<grid, CT_ROWS, 0, ctx->stream>>>(row0, nrows, rowptr, colidx, vals, X, csX,
                                                                  n_local, ghost, Y, csY, k);
    } else {
        dim3 grid((unsigned)ceil_div64(nrows, CSR_THREADS), (k + CJ - 1) / CJ);
        csr_apply_kernel<<<grid, CSR_THREADS, 0, ctx->stream>>>(row0, nrows, rowptr, colidx, vals, X, csX,
                                                              n_local, ghost, Y, csY, k);
    }
    pmc_timing_end(ctx, tslot);
    ctx->launches += 1;
    PMC_CHECK_CUDA(cudaGetLastError());
    return PMC_OK;
}

}  // namespace

extern "C" int pmc_csr_apply(pmc_ctx *ctx, int64_t nrows, const int64_t *rowptr, const int32_t *colidx,
                             const double *vals, const double *X, int64_t csX, const double *ghost,
                             double *Y, int64_t csY, int32_t k, uint32_t flags) {
    PMC_REQUIRE(ctx && nrows >= 0 && k >= 0, "bad arguments");
    if (nrows == 0 || k == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(rowptr && colidx && vals && X && Y, "NULL pointer");
    int rc = launch_csr_apply(ctx, 0, nrows, rowptr, colidx, vals, X, csX, nrows, ghost, Y, csY, k);
    if (rc) return rc;
    return pmc_after_launch(ctx, 0, flags);
}

extern "C" int pmc_csr_gram(pmc_ctx *ctx, int64_t nrows, const int64_t *rowptr, const int32_t *colidx,
                            const double *vals, const double *V, int64_t csV, int32_t kV, const double *U,
                            int64_t csU, int32_t kU, const double *ghost, double *out, int64_t ldo,
                            uint32_t flags) {
    PMC_REQUIRE(ctx && nrows >= 0 && kV >= 0 && kU >= 0, "bad arguments");
    if (kV == 0 || kU == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(out && ldo >= kV, "bad output");
    if (nrows == 0) {
        PMC_CHECK_CUDA(cudaMemset2DAsync(out, ldo * sizeof(double), 0, kV * sizeof(double), kU, ctx->stream));
        return pmc_after_launch(ctx, 0, flags);
    }
    PMC_REQUIRE(rowptr && colidx && vals && V && U, "NULL pointer");

    // Workspace layout: [ T panel (PR x kU, leading dimension PR) | Gram split partials ].
    // The panel is sized for the Gram kernel, not for L2: with >= ~30k rows per panel the split-n
    // decomposition has enough items to fill the SMs evenly (a 3k-row panel at k = 2000 gave 256 items
    // on 148 SMs, i.e. two ragged waves), while writing the panel to HBM and reading it back costs
    // 16*n*k bytes in total -- 2 % of the contraction time at k = 2000.
    const size_t panel_budget = (size_t)512 << 20;
    int64_t PR = (int64_t)(panel_budget / ((size_t)kU * 8));
    PR = (PR / 128) * 128;
    if (PR < 2048) PR = 2048;
    if (PR > nrows) PR = round_up64(nrows, 16);
    const size_t panel_bytes = (size_t)PR * kU * 8;
    const size_t panel_region = (panel_bytes + 255) & ~(size_t)255;
    if (ctx->ws == nullptr || ctx->ws_bytes < panel_region + ((size_t)64 << 20)) {
        ctx->ws_needed = panel_region + ((size_t)512 << 20);
        pmc_set_error("workspace too small: need %zu bytes, have %zu", ctx->ws_needed, ctx->ws_bytes);
        return PMC_ERR_WORKSPACE;
    }
    double *T = static_cast<double *>(ctx->ws);
    void *gws = static_cast<char *>(ctx->ws) + panel_region;
    const size_t gws_cap = ctx->ws_bytes - panel_region;
    // PMC_LOWER_ONLY: the caller knows the result is symmetric (V and U are the same panel and M is symmetric):
    // only the tiles on/below the diagonal are contracted (half the flop), the rest is mirrored
    PMC_REQUIRE(!(flags & PMC_LOWER_ONLY) || kV == kU, "PMC_LOWER_ONLY needs a square result");
    const uint32_t gflags = flags & (PMC_FORCE_SIMPLE | PMC_LOWER_ONLY);

    int first = 1;
    for (int64_t r0 = 0; r0 < nrows; r0 += PR) {
        const int64_t pr = (nrows - r0 < PR) ? nrows - r0 : PR;
        int rc = launch_csr_apply(ctx, r0, pr, rowptr, colidx, vals, U, csU, nrows, ghost, T, PR, kU);
        if (rc) return rc;
        rc = pmc_gram_impl(ctx, V + r0, csV, kV, T, PR, kU, pr, nullptr, out, ldo, gflags, gws, gws_cap,
                           first ? 0 : 1);
        if (rc) return rc;
        first = 0;
    }
    return pmc_after_launch(ctx, 0, flags);
}

extern "C" int pmc_csr_pairwise(pmc_ctx *ctx, int64_t nrows, const int64_t *rowptr, const int32_t *colidx,
                                const double *vals, const double *V, int64_t csV, const double *U, int64_t csU,
                                const double *ghost, int32_t k, double *out, uint32_t flags) {
    PMC_REQUIRE(ctx && nrows >= 0 && k >= 0, "bad arguments");
    if (k == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(out != nullptr, "out is NULL");
    if (nrows == 0) {
        PMC_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * k, ctx->stream));
        return pmc_after_launch(ctx, 0, flags);
    }
    PMC_REQUIRE(rowptr && colidx && vals && V && U, "NULL pointer");
    int64_t nb = ceil_div64(nrows, CSR_THREADS);
    if (nb > CP_MAXBLOCKS) nb = CP_MAXBLOCKS;
    const int groups = (k + CJ - 1) / CJ;
    int rc = pmc_need_ws(ctx, (size_t)nb * groups * CJ * sizeof(double));
    if (rc) return rc;
    double *partials = static_cast<double *>(ctx->ws);
    const int tslot = pmc_timing_begin(ctx, PMC_KERNEL_SPMM);
    csr_pairwise_kernel<<<dim3((unsigned)nb, groups), CSR_THREADS, 0, ctx->stream>>>(
        nrows, rowptr, colidx, vals, V, csV, U, csU, ghost, k, partials);
    pmc_timing_end(ctx, tslot);
    PMC_CHECK_CUDA(cudaGetLastError());
    csr_pairwise_reduce_kernel<<<k, 256, 0, ctx->stream>>>(partials, (int)nb, k, out);
    return pmc_after_launch(ctx, 2, flags);
}

extern "C" int pmc_gather_rows(pmc_ctx *ctx, const double *A, int64_t csA, int32_t k, int64_t n,
                               const int64_t *idx_dev, int64_t nidx, double *out, uint32_t flags) {
    PMC_REQUIRE(ctx && k >= 0 && n >= 0 && nidx >= 0, "bad arguments");
    if (k == 0 || nidx == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(A && idx_dev && out, "NULL pointer");
    const int64_t total = nidx * k;
    gather_rows_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, ctx->stream>>>(A, csA, k, idx_dev, nidx, out);
    return pmc_after_launch(ctx, 1, flags);
}
