// HBM-bound column kernels: pairwise_inner / norm2 (optionally diag-weighted), axpy, scal, copy,
// gather, dofs, amax, diag_apply.  Coalesced 128-bit streaming loads, warp-shuffle + block
// reduction, one launch per call (last-block ticket for the cross-block sum, fixed summation order
// => bitwise reproducible results).
#include "common.cuh"

namespace {

constexpr int TPB = 256;          // threads per block
constexpr int UNROLL = 4;         // independent 128-bit loads in flight per operand per thread
constexpr int CHUNK = TPB * 2 * UNROLL;  // rows per block-iteration (2048)
constexpr int MAX_COLS_PER_LAUNCH = 4096;
constexpr int PACK = 256;         // scalars / indices carried in kernel parameters per launch
constexpr int64_t L2_KEEP_BYTES = 96ll << 20;  // single vectors up to this size are kept L2-resident

struct AlphaPack { double v[PACK]; };
struct IndexPack { int64_t v[PACK]; };

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool vec_ok(const double *p, int64_t cs) { return p == nullptr || (aligned16(p) && (cs % 2) == 0); }

inline int blocks_per_col(const pmc_ctx *ctx, int64_t n, int k, int blocks_per_sm) {
    int64_t want = ceil_div64(n, CHUNK);
    int64_t cap = ((int64_t)ctx->sm_count * blocks_per_sm + k - 1) / k;
    if (cap < 1) cap = 1;
    int64_t bx = want < cap ? want : cap;
    return (int)(bx < 1 ? 1 : bx);
}

// ---------------------------------------------------------------------------------------------
// column reductions:  out[c] = sum_d a[d] * (NORM ? a[d] : b[d]) * (w ? w[d] : 1)
// ---------------------------------------------------------------------------------------------
template <bool NORM, bool WEIGHTED, bool VEC>
__global__ void __launch_bounds__(TPB)
reduce_cols_kernel(const double *__restrict__ A, int64_t csA, const double *__restrict__ B,
                   int64_t csB, const double *__restrict__ w, int64_t n,
                   double *__restrict__ partials, unsigned int *__restrict__ tickets,
                   double *__restrict__ out, int keep) {
    const int c = blockIdx.y;
    const int bx = gridDim.x;
    const double *a = A + (int64_t)c * csA;
    const double *b = NORM ? a : B + (int64_t)c * csB;
    // keep: the last operand (the vector being worked on) stays in L2, the others stream past it
    const uint64_t pol_hot = l2_policy(keep ? 2 : 0), pol_cold = l2_policy(keep ? 1 : 0);
    const uint64_t pol_a = NORM ? pol_hot : pol_cold;
    double acc[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc[u] = 0.0;

    for (int64_t base = (int64_t)blockIdx.x * CHUNK; base < n; base += (int64_t)bx * CHUNK) {
        if (VEC && base + CHUNK <= n) {
            double2 av[UNROLL], bv[UNROLL], wv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t i = base + 2 * (threadIdx.x + u * TPB);
                av[u] = ld_hint2_nc(a + i, pol_a);
                if (!NORM) bv[u] = ld_hint2_nc(b + i, pol_hot);
                if (WEIGHTED) wv[u] = ld_hint2_nc(w + i, pol_cold);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                double2 x = av[u];
                double2 y = NORM ? av[u] : bv[u];
                if (WEIGHTED) { y.x *= wv[u].x; y.y *= wv[u].y; }
                acc[u] = fma(x.x, y.x, acc[u]);
                acc[u] = fma(x.y, y.y, acc[u]);
            }
        } else {
            const int64_t end = (base + CHUNK < n) ? base + CHUNK : n;
            for (int64_t i = base + threadIdx.x; i < end; i += TPB) {
                double x = a[i];
                double y = NORM ? x : b[i];
                if (WEIGHTED) y *= w[i];
                acc[0] = fma(x, y, acc[0]);
            }
        }
    }
    double s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    s = warp_sum(s);
    __shared__ double wsum[TPB / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wsum[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < TPB / 32; ++i) t += wsum[i];
        if (bx == 1) {
            out[c] = t;
            is_last = false;
        } else {
            partials[(int64_t)c * bx + blockIdx.x] = t;
            __threadfence();
            unsigned int prev = atomicAdd(&tickets[c], 1u);
            is_last = (prev == (unsigned int)bx - 1);
        }
    }
    __syncthreads();
    if (is_last && warp == 0) {
        __threadfence();
        double t = 0.0;
        const volatile double *p = partials + (int64_t)c * bx;
        for (int i = lane; i < bx; i += 32) t += p[i];
        t = warp_sum(t);
        if (lane == 0) {
            out[c] = t;
            tickets[c] = 0;  // self-reset for the next launch
        }
    }
}

template <bool NORM>
int launch_reduce(pmc_ctx *ctx, const double *A, int64_t csA, const double *B, int64_t csB, int32_t k,
                  int64_t n, const double *w, double *out, uint32_t flags, double *partials_buf,
                  int partials_cap) {
    int launches = 0;
    for (int c0 = 0; c0 < k; c0 += MAX_COLS_PER_LAUNCH) {
        int kc = (k - c0 < MAX_COLS_PER_LAUNCH) ? k - c0 : MAX_COLS_PER_LAUNCH;
        int bx = blocks_per_col(ctx, n, kc, 8);
        while ((int64_t)bx * kc > partials_cap && bx > 1) bx = (bx + 1) / 2;
        const double *a = A + (int64_t)c0 * csA;
        const double *b = NORM ? nullptr : B + (int64_t)c0 * csB;
        bool vec = vec_ok(a, csA) && (NORM || vec_ok(b, csB)) && vec_ok(w, 0);
        // one hot vector: a single pair (k == 1), or one column broadcast against a wide panel (csB == 0: the
        // block projection r = Q^T W v of classical Gram-Schmidt) -- the panel streams past it evict_first
        const bool one_hot = k == 1 || (!NORM && csB == 0 && csA != 0);
        const int keep = (ctx->l2_keep && one_hot && n * 8 <= L2_KEEP_BYTES) ? 1 : 0;
        dim3 grid(bx, kc);
#define PMC_LAUNCH_RED(WT, VC)                                                                   \
    reduce_cols_kernel<NORM, WT, VC><<<grid, TPB, 0, ctx->stream>>>(a, csA, b, csB, w, n,          \
                                                                   partials_buf, ctx->tickets,    \
                                                                   out + c0, keep)
        if (w) { if (vec) PMC_LAUNCH_RED(true, true); else PMC_LAUNCH_RED(true, false); }
        else   { if (vec) PMC_LAUNCH_RED(false, true); else PMC_LAUNCH_RED(false, false); }
#undef PMC_LAUNCH_RED
        ++launches;
    }
    return pmc_after_launch(ctx, launches, flags);
}

// ---------------------------------------------------------------------------------------------
// fused Gram-Schmidt step:  y += alpha * x ;  out = sum_d q[d] * w[d] * y_new[d]   (q == y: norm2)
// One pass instead of an axpy launch followed by a reduction launch: the host layer defers the
// axpy of `v.axpy(-p, q_j)` until the next `q_{j+1}.pairwise_inner(v)` / `v.norm()` arrives.
// Same arithmetic and the same summation order as the two separate kernels (bitwise identical).
// ---------------------------------------------------------------------------------------------
template <bool NORM, bool WEIGHTED>
__global__ void __launch_bounds__(TPB)
axpy_dot_kernel(double *__restrict__ y, const double *__restrict__ x, double alpha,
                const double *__restrict__ q, const double *__restrict__ w, int64_t n,
                double *__restrict__ partials, unsigned int *__restrict__ tickets,
                double *__restrict__ out, int keep) {
    const int bx = gridDim.x;
    const uint64_t pol_hot = l2_policy(keep ? 2 : 0), pol_cold = l2_policy(keep ? 1 : 0);
    double acc[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc[u] = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * CHUNK; base < n; base += (int64_t)bx * CHUNK) {
        if (base + CHUNK <= n) {
            double2 xv[UNROLL], yv[UNROLL], qv[UNROLL], wv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t i = base + 2 * (threadIdx.x + u * TPB);
                xv[u] = ld_hint2_nc(x + i, pol_cold);
                yv[u] = ld_hint2(y + i, pol_hot);
                if (!NORM) qv[u] = ld_hint2_nc(q + i, pol_cold);
                if (WEIGHTED) wv[u] = ld_hint2_nc(w + i, pol_cold);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t i = base + 2 * (threadIdx.x + u * TPB);
                double2 r;
                r.x = fma(alpha, xv[u].x, yv[u].x);
                r.y = fma(alpha, xv[u].y, yv[u].y);
                st_hint2(y + i, r, pol_hot);
                double2 a = NORM ? r : qv[u];
                double2 b = r;
                if (WEIGHTED) { b.x *= wv[u].x; b.y *= wv[u].y; }
                acc[u] = fma(a.x, b.x, acc[u]);
                acc[u] = fma(a.y, b.y, acc[u]);
            }
        } else {
            const int64_t end = n;
            for (int64_t i = base + threadIdx.x; i < end; i += TPB) {
                const double r = fma(alpha, x[i], y[i]);
                y[i] = r;
                const double a = NORM ? r : q[i];
                double b = r;
                if (WEIGHTED) b *= w[i];
                acc[0] = fma(a, b, acc[0]);
            }
        }
    }
    double s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    s = warp_sum(s);
    __shared__ double wsum[TPB / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wsum[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < TPB / 32; ++i) t += wsum[i];
        if (bx == 1) {
            out[0] = t;
            is_last = false;
        } else {
            partials[blockIdx.x] = t;
            __threadfence();
            is_last = (atomicAdd(&tickets[0], 1u) == (unsigned int)bx - 1);
        }
    }
    __syncthreads();
    if (is_last && warp == 0) {
        __threadfence();
        double t = 0.0;
        const volatile double *p = partials;
        for (int i = lane; i < bx; i += 32) t += p[i];
        t = warp_sum(t);
        if (lane == 0) {
            out[0] = t;
            tickets[0] = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// elementwise column kernels
// ---------------------------------------------------------------------------------------------
enum { OP_AXPY = 0, OP_SCAL = 1, OP_COPY = 2, OP_DIAG = 3 };

template <int OP, bool VEC>
__global__ void __launch_bounds__(TPB)
elementwise_kernel(double *__restrict__ Y, int64_t csY, const double *X, int64_t csX,
                   const double *__restrict__ w, int64_t n, AlphaPack alpha, int alpha_stride, int keep) {
    const uint64_t pol_hot = l2_policy(keep ? 2 : 0), pol_cold = l2_policy(keep ? 1 : 0);
    const int c = blockIdx.y;
    double *y = Y + (int64_t)c * csY;
    const double *x = (OP == OP_SCAL) ? nullptr : X + (int64_t)c * csX;
    const double al = alpha.v[c * alpha_stride];
    const int bx = gridDim.x;
    for (int64_t base = (int64_t)blockIdx.x * CHUNK; base < n; base += (int64_t)bx * CHUNK) {
        if (VEC && base + CHUNK <= n) {
            double2 xv[UNROLL], yv[UNROLL], wv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t i = base + 2 * (threadIdx.x + u * TPB);
                if (OP != OP_SCAL) xv[u] = ld_hint2_nc(x + i, pol_cold);
                if (OP == OP_AXPY || OP == OP_SCAL) yv[u] = ld_hint2(y + i, pol_hot);
                if (OP == OP_DIAG) wv[u] = ld_hint2_nc(w + i, pol_cold);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t i = base + 2 * (threadIdx.x + u * TPB);
                double2 r;
                if (OP == OP_AXPY) { r.x = fma(al, xv[u].x, yv[u].x); r.y = fma(al, xv[u].y, yv[u].y); }
                if (OP == OP_SCAL) { r.x = al * yv[u].x; r.y = al * yv[u].y; }
                if (OP == OP_COPY) { r = xv[u]; }
                if (OP == OP_DIAG) { r.x = wv[u].x * xv[u].x; r.y = wv[u].y * xv[u].y; }
                st_hint2(y + i, r, pol_hot);
            }
        } else {
            const int64_t end = (base + CHUNK < n) ? base + CHUNK : n;
            for (int64_t i = base + threadIdx.x; i < end; i += TPB) {
                double r;
                if (OP == OP_AXPY) r = fma(al, x[i], y[i]);
                if (OP == OP_SCAL) r = al * y[i];
                if (OP == OP_COPY) r = x[i];
                if (OP == OP_DIAG) r = w[i] * x[i];
                y[i] = r;
            }
        }
    }
}

template <int OP>
int launch_elementwise(pmc_ctx *ctx, double *Y, int64_t csY, const double *X, int64_t csX,
                       const double *w, int32_t k, int64_t n, const double *alpha, int32_t n_alpha,
                       uint32_t flags) {
    int launches = 0;
    for (int c0 = 0; c0 < k; c0 += PACK) {
        int kc = (k - c0 < PACK) ? k - c0 : PACK;
        AlphaPack pack;
        int stride = 0;
        if (alpha) {
            if (n_alpha == 1) {
                pack.v[0] = alpha[0];
            } else {
                stride = 1;
                memcpy(pack.v, alpha + c0, sizeof(double) * kc);
            }
        } else {
            pack.v[0] = 1.0;
        }
        double *y = Y + (int64_t)c0 * csY;
        const double *x = X ? X + (int64_t)c0 * csX : nullptr;
        bool vec = vec_ok(y, csY) && vec_ok(x, csX) && vec_ok(w, 0);
        int bx = blocks_per_col(ctx, n, kc, 8);
        dim3 grid(bx, kc);
        const int keep = (ctx->l2_keep && (OP == OP_AXPY || OP == OP_SCAL) && k == 1 && n * 8 <= L2_KEEP_BYTES) ? 1 : 0;
        if (vec)
            elementwise_kernel<OP, true><<<grid, TPB, 0, ctx->stream>>>(y, csY, x, csX, w, n, pack, stride, keep);
        else
            elementwise_kernel<OP, false><<<grid, TPB, 0, ctx->stream>>>(y, csY, x, csX, w, n, pack, stride, keep);
        ++launches;
    }
    return pmc_after_launch(ctx, launches, flags);
}

// gather of arbitrary columns
template <bool VEC>
__global__ void __launch_bounds__(TPB)
gather_cols_kernel(double *__restrict__ D, int64_t csD, const double *__restrict__ base, int64_t ld,
                   int64_t n, IndexPack idx) {
    const int c = blockIdx.y;
    double *y = D + (int64_t)c * csD;
    const double *x = base + idx.v[c] * ld;
    const int bx = gridDim.x;
    for (int64_t b0 = (int64_t)blockIdx.x * CHUNK; b0 < n; b0 += (int64_t)bx * CHUNK) {
        if (VEC && b0 + CHUNK <= n) {
            double2 xv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) xv[u] = ld_stream2(x + b0 + 2 * (threadIdx.x + u * TPB));
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                *reinterpret_cast<double2 *>(y + b0 + 2 * (threadIdx.x + u * TPB)) = xv[u];
        } else {
            const int64_t end = (b0 + CHUNK < n) ? b0 + CHUNK : n;
            for (int64_t i = b0 + threadIdx.x; i < end; i += TPB) y[i] = x[i];
        }
    }
}

// In-place compaction after `del A[...]`: columns [first, first + count) move `shift` columns down.  Source and
// destination ranges overlap when count > shift, so every thread owns a set of ROWS and walks the columns in
// increasing order: a column is read (by the same thread) before the move of a later column overwrites it, and
// rows never interact.  One launch per run of kept columns whatever its length.
__global__ void __launch_bounds__(TPB)
shift_cols_kernel(double *__restrict__ A, int64_t cs, int64_t first, int64_t count, int64_t shift, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * TPB * 2;
    for (int64_t i = ((int64_t)blockIdx.x * TPB + threadIdx.x) * 2; i < n; i += stride) {
        if (i + 1 < n) {
            for (int64_t c = first; c < first + count; ++c) {
                const double2 v = *reinterpret_cast<const double2 *>(A + c * cs + i);
                *reinterpret_cast<double2 *>(A + (c - shift) * cs + i) = v;
            }
        } else {
            for (int64_t c = first; c < first + count; ++c) A[(c - shift) * cs + i] = A[c * cs + i];
        }
    }
}

// After a cross-GPU all-reduce the two halves of a symmetric k x k result may differ in the last bit (the
// collective sums different chunks in different rank orders): copy the lower triangle over the upper one so that
// the Gramian stays bitwise symmetric, like the reference's syrk result.
__global__ void mirror_lower_kernel(double *__restrict__ G, int64_t k, int64_t ld) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= k * k) return;
    const int64_t i = e % k, j = e / k;
    if (i > j) G[j + i * ld] = G[i + j * ld];
}

__global__ void dofs_kernel(const double *__restrict__ A, int64_t csA, int k, IndexPack dof, int nd,
                            int ndofs_total, int dof0, double *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nd * k) return;
    const int i = t % nd, c = t / nd;
    out[(int64_t)c * ndofs_total + dof0 + i] = A[dof.v[i] + (int64_t)c * csA];
}

// abs-argmax with first-index tie-break; NaN counts as maximal (np.argmax semantics)
__device__ __forceinline__ bool amax_better(double v, int64_t i, double bv, int64_t bi) {
    const bool vn = isnan(v), bn = isnan(bv);
    if (vn != bn) return vn;
    if (!vn && v != bv) return v > bv;
    return i < bi;
}

template <bool VEC>
__global__ void __launch_bounds__(TPB)
amax_kernel(const double *__restrict__ A, int64_t csA, int64_t n, double *__restrict__ pval,
            int64_t *__restrict__ pidx, unsigned int *__restrict__ tickets,
            int64_t *__restrict__ out_idx, double *__restrict__ out_val) {
    const int c = blockIdx.y, bx = gridDim.x;
    const double *a = A + (int64_t)c * csA;
    double bv = -1.0;
    int64_t bi = INT64_MAX;
    // (amax_better is a total order on (value, index): the result does not depend on the order of the comparisons)
    for (int64_t base = (int64_t)blockIdx.x * CHUNK; base < n; base += (int64_t)bx * CHUNK) {
        if (VEC && base + CHUNK <= n) {
            double2 xv[UNROLL];                       // UNROLL independent 128-bit loads in flight
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) xv[u] = ld_stream2(a + base + 2 * (threadIdx.x + u * TPB));
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t i = base + 2 * (threadIdx.x + u * TPB);
                const double v0 = fabs(xv[u].x), v1 = fabs(xv[u].y);
                if (amax_better(v0, i, bv, bi)) { bv = v0; bi = i; }
                if (amax_better(v1, i + 1, bv, bi)) { bv = v1; bi = i + 1; }
            }
        } else {
            const int64_t end = (base + CHUNK < n) ? base + CHUNK : n;
            for (int64_t i = base + threadIdx.x; i < end; i += TPB) {
                const double v = fabs(a[i]);
                if (amax_better(v, i, bv, bi)) { bv = v; bi = i; }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (amax_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    __shared__ double sv[TPB / 32];
    __shared__ int64_t si[TPB / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < TPB / 32; ++i)
            if (amax_better(sv[i], si[i], bv, bi)) { bv = sv[i]; bi = si[i]; }
        if (bx == 1) {
            out_idx[c] = bi; out_val[c] = bv; is_last = false;
        } else {
            pval[(int64_t)c * bx + blockIdx.x] = bv;
            pidx[(int64_t)c * bx + blockIdx.x] = bi;
            __threadfence();
            is_last = (atomicAdd(&tickets[c], 1u) == (unsigned int)bx - 1);
        }
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        const volatile double *pv = pval + (int64_t)c * bx;
        const volatile int64_t *pi = pidx + (int64_t)c * bx;
        bv = pv[0]; bi = pi[0];
        for (int i = 1; i < bx; ++i)
            if (amax_better(pv[i], pi[i], bv, bi)) { bv = pv[i]; bi = pi[i]; }
        out_idx[c] = bi; out_val[c] = bv;
        tickets[c] = 0;
    }
}

}  // namespace

static int get_partials(pmc_ctx *ctx, double **buf, int *cap) {
    *buf = ctx->partials;
    *cap = ctx->partials_cap;
    return PMC_OK;
}

extern "C" int pmc_pairwise_inner(pmc_ctx *ctx, const double *A, int64_t csA, const double *B,
                                  int64_t csB, int32_t k, int64_t n, const double *w, double *out,
                                  uint32_t flags) {
    PMC_REQUIRE(ctx && k >= 0 && n >= 0, "bad arguments");
    if (k == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(out != nullptr, "out is NULL");
    if (n == 0) {
        PMC_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * k, ctx->stream));
        return pmc_after_launch(ctx, 0, flags);
    }
    PMC_REQUIRE(A && B, "NULL panel");
    double *pb; int cap;
    int rc = get_partials(ctx, &pb, &cap);
    if (rc) return rc;
    return launch_reduce<false>(ctx, A, csA, B, csB, k, n, w, out, flags, pb, cap);
}

extern "C" int pmc_norm2(pmc_ctx *ctx, const double *A, int64_t csA, int32_t k, int64_t n,
                         const double *w, double *out, uint32_t flags) {
    PMC_REQUIRE(ctx && k >= 0 && n >= 0, "bad arguments");
    if (k == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(out != nullptr, "out is NULL");
    if (n == 0) {
        PMC_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * k, ctx->stream));
        return pmc_after_launch(ctx, 0, flags);
    }
    PMC_REQUIRE(A, "NULL panel");
    double *pb; int cap;
    int rc = get_partials(ctx, &pb, &cap);
    if (rc) return rc;
    return launch_reduce<true>(ctx, A, csA, nullptr, 0, k, n, w, out, flags, pb, cap);
}

extern "C" int pmc_axpy(pmc_ctx *ctx, double *Y, int64_t csY, const double *X, int64_t csX, int32_t k,
                        int64_t n, const double *alpha, int32_t n_alpha, uint32_t flags) {
    PMC_REQUIRE(ctx && k >= 0 && n >= 0, "bad arguments");
    PMC_REQUIRE(alpha && (n_alpha == 1 || n_alpha == k), "alpha must hold 1 or k values");
    if (k == 0 || n == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(Y && X, "NULL panel");
    return launch_elementwise<OP_AXPY>(ctx, Y, csY, X, csX, nullptr, k, n, alpha, n_alpha, flags);
}

extern "C" int pmc_scal(pmc_ctx *ctx, double *Y, int64_t csY, int32_t k, int64_t n,
                        const double *alpha, int32_t n_alpha, uint32_t flags) {
    PMC_REQUIRE(ctx && k >= 0 && n >= 0, "bad arguments");
    PMC_REQUIRE(alpha && (n_alpha == 1 || n_alpha == k), "alpha must hold 1 or k values");
    if (k == 0 || n == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(Y, "NULL panel");
    return launch_elementwise<OP_SCAL>(ctx, Y, csY, nullptr, 0, nullptr, k, n, alpha, n_alpha, flags);
}

extern "C" int pmc_copy(pmc_ctx *ctx, double *D, int64_t csD, const double *S, int64_t csS, int32_t k,
                        int64_t n, uint32_t flags) {
    PMC_REQUIRE(ctx && k >= 0 && n >= 0, "bad arguments");
    if (k == 0 || n == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(D && S, "NULL panel");
    return launch_elementwise<OP_COPY>(ctx, D, csD, S, csS, nullptr, k, n, nullptr, 1, flags);
}

extern "C" int pmc_diag_apply(pmc_ctx *ctx, const double *w, const double *X, int64_t csX, double *Y,
                              int64_t csY, int32_t k, int64_t n, uint32_t flags) {
    PMC_REQUIRE(ctx && k >= 0 && n >= 0, "bad arguments");
    if (k == 0 || n == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(w && X && Y, "NULL pointer");
    return launch_elementwise<OP_DIAG>(ctx, Y, csY, X, csX, w, k, n, nullptr, 1, flags);
}

extern "C" int pmc_gather_cols(pmc_ctx *ctx, double *D, int64_t csD, const double *base, int64_t ld,
                               const int64_t *idx, int32_t k, int64_t n, uint32_t flags) {
    PMC_REQUIRE(ctx && k >= 0 && n >= 0, "bad arguments");
    if (k == 0 || n == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(D && base && idx, "NULL pointer");
    int launches = 0;
    for (int c0 = 0; c0 < k; c0 += PACK) {
        int kc = (k - c0 < PACK) ? k - c0 : PACK;
        IndexPack pack;
        memcpy(pack.v, idx + c0, sizeof(int64_t) * kc);
        double *d = D + (int64_t)c0 * csD;
        bool vec = vec_ok(d, csD) && vec_ok(base, ld);
        dim3 grid(blocks_per_col(ctx, n, kc, 8), kc);
        if (vec) gather_cols_kernel<true><<<grid, TPB, 0, ctx->stream>>>(d, csD, base, ld, n, pack);
        else gather_cols_kernel<false><<<grid, TPB, 0, ctx->stream>>>(d, csD, base, ld, n, pack);
        ++launches;
    }
    return pmc_after_launch(ctx, launches, flags);
}

extern "C" int pmc_mirror_lower(pmc_ctx *ctx, double *G, int64_t k, int64_t ld, uint32_t flags) {
    PMC_REQUIRE(ctx && k >= 0 && ld >= k, "bad arguments");
    if (k <= 1) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(G != nullptr, "G is NULL");
    mirror_lower_kernel<<<(unsigned)ceil_div64(k * k, 256), 256, 0, ctx->stream>>>(G, k, ld);
    return pmc_after_launch(ctx, 1, flags);
}

extern "C" int pmc_shift_cols(pmc_ctx *ctx, double *A, int64_t cs, int64_t first, int64_t count, int64_t shift,
                              int64_t n, uint32_t flags) {
    PMC_REQUIRE(ctx && first >= 0 && count >= 0 && shift >= 0 && shift <= first && n >= 0, "bad arguments");
    if (count == 0 || shift == 0 || n == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(A && cs >= n && (cs % 2) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0,
                "pmc_shift_cols needs a 16-byte aligned panel with an even column stride >= n");
    int64_t blocks = ceil_div64(n, 2 * TPB);
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    shift_cols_kernel<<<(unsigned)blocks, TPB, 0, ctx->stream>>>(A, cs, first, count, shift, n);
    return pmc_after_launch(ctx, 1, flags);
}

extern "C" int pmc_dofs(pmc_ctx *ctx, const double *A, int64_t csA, int32_t k, int64_t n,
                        const int64_t *dof, int32_t ndofs, double *out, uint32_t flags) {
    PMC_REQUIRE(ctx && k >= 0 && n >= 0 && ndofs >= 0, "bad arguments");
    if (k == 0 || ndofs == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(A && dof && out, "NULL pointer");
    for (int i = 0; i < ndofs; ++i) PMC_REQUIRE(dof[i] >= 0 && dof[i] < n, "dof index out of range");
    int launches = 0;
    for (int d0 = 0; d0 < ndofs; d0 += PACK) {
        int nd = (ndofs - d0 < PACK) ? ndofs - d0 : PACK;
        IndexPack pack;
        memcpy(pack.v, dof + d0, sizeof(int64_t) * nd);
        int total = nd * k;
        dofs_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(A, csA, k, pack, nd, ndofs, d0, out);
        ++launches;
    }
    return pmc_after_launch(ctx, launches, flags);
}

extern "C" int pmc_amax(pmc_ctx *ctx, const double *A, int64_t csA, int32_t k, int64_t n,
                        int64_t *out_idx, double *out_val, uint32_t flags) {
    PMC_REQUIRE(ctx && k >= 0 && n >= 1, "amax needs n >= 1");
    if (k == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(A && out_idx && out_val, "NULL pointer");
    double *pb; int cap;
    int rc = get_partials(ctx, &pb, &cap);
    if (rc) return rc;
    int64_t *pi = reinterpret_cast<int64_t *>(pb + cap);
    int launches = 0;
    for (int c0 = 0; c0 < k; c0 += MAX_COLS_PER_LAUNCH) {
        int kc = (k - c0 < MAX_COLS_PER_LAUNCH) ? k - c0 : MAX_COLS_PER_LAUNCH;
        int bx = blocks_per_col(ctx, n, kc, 8);
        while ((int64_t)bx * kc > cap && bx > 1) bx = (bx + 1) / 2;
        dim3 grid(bx, kc);
        if (vec_ok(A, csA))
            amax_kernel<true><<<grid, TPB, 0, ctx->stream>>>(A + (int64_t)c0 * csA, csA, n, pb, pi, ctx->tickets,
                                                             out_idx + c0, out_val + c0);
        else
            amax_kernel<false><<<grid, TPB, 0, ctx->stream>>>(A + (int64_t)c0 * csA, csA, n, pb, pi, ctx->tickets,
                                                              out_idx + c0, out_val + c0);
        ++launches;
    }
    return pmc_after_launch(ctx, launches, flags);
}

extern "C" int pmc_axpy_dot(pmc_ctx *ctx, double *Y, const double *X, const double *alpha, const double *Q,
                            int64_t n, const double *w, double *out, uint32_t flags) {
    PMC_REQUIRE(ctx && n >= 0 && alpha && out, "bad arguments");
    if (n == 0) {
        PMC_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double), ctx->stream));
        return pmc_after_launch(ctx, 0, flags);
    }
    PMC_REQUIRE(Y && X, "NULL vector");
    const bool vec = aligned16(Y) && aligned16(X) && (!Q || aligned16(Q)) && (!w || aligned16(w));
    if (!vec) {   // misaligned vectors: same result from the two separate kernels
        int rc = pmc_axpy(ctx, Y, 0, X, 0, 1, n, alpha, 1, 0);
        if (rc) return rc;
        return Q ? pmc_pairwise_inner(ctx, Q, 0, Y, 0, 1, n, w, out, flags) : pmc_norm2(ctx, Y, 0, 1, n, w, out, flags);
    }
    const int bx = blocks_per_col(ctx, n, 1, 8);
    const int keep = (ctx->l2_keep && n * 8 <= L2_KEEP_BYTES) ? 1 : 0;
#define PMC_LAUNCH_AD(NM, WT)                                                                            \
    axpy_dot_kernel<NM, WT><<<bx, TPB, 0, ctx->stream>>>(Y, X, alpha[0], Q, w, n, ctx->partials, ctx->tickets, \
                                                        out, keep)
    if (Q) { if (w) PMC_LAUNCH_AD(false, true); else PMC_LAUNCH_AD(false, false); }
    else   { if (w) PMC_LAUNCH_AD(true, true); else PMC_LAUNCH_AD(true, false); }
#undef PMC_LAUNCH_AD
    return pmc_after_launch(ctx, 1, flags);
}
