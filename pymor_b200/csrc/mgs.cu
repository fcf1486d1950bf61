// Gram-Schmidt sweep: orthogonalise ONE vector against a panel of orthonormal vectors in ONE launch.
//
//   for j = 0 .. nq-1:   r[j] = <q_j, v>_w  (summed over all row shards / GPUs);   v -= r[j] * q_j
//   norm2 = <v, v>_w
//
// This is the inner loop of gram_schmidt (src/pymor/algorithms/gram_schmidt.py:84-93).  Driven from the
// host it costs one launch + one stream synchronisation (+ one cross-rank sum) per pair (q_j, v):
// 74 us per pair-step at 10M rows on one B200 and a 35-40 us latency floor at 8 GPUs, where the rows
// of a shard stream through in 3 us (profiles/r01_summary.md).  Here the whole sweep is one persistent
// cooperative kernel, one CTA per SM:
//
//   * every CTA owns a fixed set of rows, so the update of v needs no synchronisation at all; the
//     deferred update  v -= r[j-1] q_{j-1}  is merged into the pass that accumulates <q_j, v>  (one pass
//     over q_{j-1}, q_j, w from HBM and v from L2 per step; with a pre-weighted panel z_j = w * q_j, kept
//     by the caller next to Q, the dots read z_j and the w stream disappears: two HBM streams per step);
//   * the only grid-wide exchange per step is the sum of the CTA partials.  Each CTA posts its partial
//     as a flagged packet (two 8-byte words, each carrying 32 value bits and the 32-bit step number, so
//     a reader never sees a torn value and no fence or atomic is needed) and every CTA polls all packets:
//     barrier and reduction in one L2 round trip.  Fixed summation order => all CTAs get the same bits;
//   * with several GPUs, CTA 0 posts the GPU's sum the same way into the inbox of every peer over
//     NVLink (peer memory opened with cudaIpc, pmc_peer_*), and every CTA of every GPU polls its local
//     inbox and adds the per-rank sums in rank order: compute and "all-reduce" are one kernel, all ranks
//     hold bitwise identical coefficients, and the host is involved once per sweep instead of once per
//     pair.  Packets are double-buffered by the parity of the GLOBAL step number (steps are counted across
//     sweeps); a rank can only be one step ahead of the slowest (it needs everybody's packet of step s to
//     start step s+1), so two buffers suffice (tests/test_mgs_sweep_model.py simulates the protocol under
//     random schedules -- it caught the per-sweep parity this started with).
//
// Results: out[0..nq-1] = r, out[nq] = norm2, out[nq+1] = status (0 ok, 1 a wait timed out).
// Summation order differs from the pair-by-pair kernels (blas1.cu), so results agree to rounding, not
// bitwise.  Opt-in until it has run on a B200 (written after the round-1 GPU budget was spent).
#include "common.cuh"

namespace {

constexpr int S_TPB = 1024;                     // one CTA per SM
constexpr int S_UNROLL = 2;                     // 128-bit loads in flight per operand and thread
constexpr int S_CHUNK = S_TPB * 2 * S_UNROLL;   // rows per CTA iteration (4096)
constexpr int64_t S_KEEP_BYTES = 96ll << 20;
constexpr unsigned long long S_TIMEOUT_NS = 5000000000ull;

struct SweepParams {
    const double *Q;
    int64_t csQ;
    int nq;
    const double *Z;            // optional pre-weighted panel z_j = w * q_j (then the dot needs no w stream)
    int64_t csZ;
    double *v;
    const double *w;
    int64_t n;
    double *out;
    uint64_t *cta_box;          // [2 parities][grid][2 words]
    unsigned int *abort_flag;
    uint32_t seq_base;          // step j of this sweep carries flag seq_base + j + 1
    int rank, size;
    uint64_t *inbox;            // local: [size][2 parities][2 words]
    uint64_t *peer_inbox[PMC_MAX_PEERS];
    int keep_v, keep_w;
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// flagged packet: word0 = value bits 31..0 | flag << 32, word1 = value bits 63..32 | flag << 32
__device__ __forceinline__ void ll_post(uint64_t *slot, double value, uint32_t flag) {
    const uint64_t bits = (uint64_t)__double_as_longlong(value);
    const uint64_t f = (uint64_t)flag << 32;
    const uint64_t w0 = (bits & 0xffffffffull) | f, w1 = (bits >> 32) | f;
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(w0) : "memory");
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(slot + 1), "l"(w1) : "memory");
}

__device__ __forceinline__ bool ll_try(const uint64_t *slot, uint32_t flag, double &value) {
    uint64_t w0, w1;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w0) : "l"(slot) : "memory");
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w1) : "l"(slot + 1) : "memory");
    if ((uint32_t)(w0 >> 32) != flag || (uint32_t)(w1 >> 32) != flag) return false;
    value = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    return true;
}

// spin until the packet of step `flag` is there; gives up (value 0, abort flag set) after S_TIMEOUT_NS or
// as soon as anybody else has given up, so that a protocol error cannot hang the GPU
__device__ __forceinline__ double ll_wait(const uint64_t *slot, uint32_t flag, unsigned int *abort_flag) {
    double value = 0.0;
    if (ll_try(slot, flag, value)) return value;
    const unsigned long long t0 = global_ns();
    unsigned int spins = 0;
    while (!ll_try(slot, flag, value)) {
        if ((++spins & 1023u) == 0) {
            if (*reinterpret_cast<volatile unsigned int *>(abort_flag) != 0u) return 0.0;
            if (global_ns() - t0 > S_TIMEOUT_NS) {
                atomicExch(abort_flag, 1u);
                return 0.0;
            }
        }
    }
    return value;
}

// One pass over this CTA's rows.  HAVE_X: v += alpha * x first (the deferred update of the previous
// step; v is written back).  NORM: accumulate v*w*v, else q*w*v.
template <bool HAVE_X, bool NORM, bool WEIGHTED>
__device__ __forceinline__ double sweep_pass(double *__restrict__ v, const double *__restrict__ x, double alpha,
                                             const double *__restrict__ q, const double *__restrict__ w, int64_t n,
                                             uint64_t pol_v, uint64_t pol_w, uint64_t pol_q) {
    double acc[S_UNROLL];
#pragma unroll
    for (int u = 0; u < S_UNROLL; ++u) acc[u] = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * S_CHUNK; base < n; base += (int64_t)gridDim.x * S_CHUNK) {
        if (base + S_CHUNK <= n) {
            double2 xv[S_UNROLL], yv[S_UNROLL], qv[S_UNROLL], wv[S_UNROLL];
#pragma unroll
            for (int u = 0; u < S_UNROLL; ++u) {
                const int64_t i = base + 2 * (threadIdx.x + u * S_TPB);
                if (HAVE_X) xv[u] = ld_hint2_nc(x + i, pol_q);
                yv[u] = ld_hint2(v + i, pol_v);
                if (!NORM) qv[u] = ld_hint2_nc(q + i, pol_q);
                if (WEIGHTED) wv[u] = ld_hint2_nc(w + i, pol_w);
            }
#pragma unroll
            for (int u = 0; u < S_UNROLL; ++u) {
                const int64_t i = base + 2 * (threadIdx.x + u * S_TPB);
                double2 r = yv[u];
                if (HAVE_X) {
                    r.x = fma(alpha, xv[u].x, yv[u].x);
                    r.y = fma(alpha, xv[u].y, yv[u].y);
                    st_hint2(v + i, r, pol_v);
                }
                const double2 a = NORM ? r : qv[u];
                double2 b = r;
                if (WEIGHTED) { b.x *= wv[u].x; b.y *= wv[u].y; }
                acc[u] = fma(a.x, b.x, acc[u]);
                acc[u] = fma(a.y, b.y, acc[u]);
            }
        } else {
            for (int64_t i = base + threadIdx.x; i < n; i += S_TPB) {
                double r = v[i];
                if (HAVE_X) {
                    r = fma(alpha, x[i], r);
                    v[i] = r;
                }
                const double a = NORM ? r : q[i];
                double b = r;
                if (WEIGHTED) b *= w[i];
                acc[0] = fma(a, b, acc[0]);
            }
        }
    }
    double s = acc[0];
#pragma unroll
    for (int u = 1; u < S_UNROLL; ++u) s += acc[u];
    return s;
}

// WMODE 0: Euclidean; 1: diagonal weight applied in every pass (q_{j-1}, q_j, w stream from HBM);
// 2: pre-weighted panel Z (z_j = w * q_j, written once when q_j became final): the dots read z_j instead of
// q_j and no w -- two HBM streams per step instead of three; only the closing norm pass reads w.
template <int WMODE>
__global__ void __launch_bounds__(S_TPB, 1) mgs_sweep_kernel(const SweepParams p) {
    constexpr bool W_DOT = WMODE == 1, W_NORM = WMODE != 0;
    __shared__ double wsum[S_TPB / 32];
    __shared__ double s_coeff;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = gridDim.x;
    const uint64_t pol_v = l2_policy(p.keep_v ? 2 : 0), pol_w = l2_policy(p.keep_w ? 2 : (p.keep_v ? 1 : 0));
    const uint64_t pol_q = l2_policy(p.keep_v ? 1 : 0);

    double coeff = 0.0;                       // r[j-1]: the update still owed to v
    for (int j = 0; j <= p.nq; ++j) {
        const double *x = p.Q + (int64_t)(j - 1) * p.csQ;      // q_{j-1} (unused for j == 0)
        // dot operand of step j: q_j, or z_j = w * q_j from the pre-weighted panel (unused for j == nq)
        const double *q = (WMODE == 2) ? p.Z + (int64_t)j * p.csZ : p.Q + (int64_t)j * p.csQ;
        double s;
        if (j == 0) {
            s = (p.nq == 0) ? sweep_pass<false, true, W_NORM>(p.v, x, 0.0, q, p.w, p.n, pol_v, pol_w, pol_q)
                            : sweep_pass<false, false, W_DOT>(p.v, x, 0.0, q, p.w, p.n, pol_v, pol_w, pol_q);
        } else if (j < p.nq) {
            s = sweep_pass<true, false, W_DOT>(p.v, x, -coeff, q, p.w, p.n, pol_v, pol_w, pol_q);
        } else {
            s = sweep_pass<true, true, W_NORM>(p.v, x, -coeff, q, p.w, p.n, pol_v, pol_w, pol_q);
        }
        // CTA partial
        s = warp_sum(s);
        if (lane == 0) wsum[warp] = s;
        __syncthreads();
        if (warp == 0) {
            double t = wsum[lane];             // S_TPB / 32 == 32 warps
            t = warp_sum(t);
            const uint32_t flag = p.seq_base + (uint32_t)j + 1u;
            const int par = (int)(flag & 1u);    // parity of the GLOBAL step number: consecutive steps of
                                                 // consecutive sweeps must not share a buffer either
            uint64_t *box = p.cta_box + (size_t)par * G * 2;
            if (lane == 0) ll_post(box + 2 * blockIdx.x, t, flag);
            // grid sum: every CTA reads every CTA's packet (barrier + reduction in one round trip)
            double g = 0.0;
            for (int i = lane; i < G; i += 32) g += ll_wait(box + 2 * i, flag, p.abort_flag);
            g = warp_sum(g);
            if (p.size > 1) {
                // all-reduce over the GPUs: CTA 0 posts this GPU's sum into every rank's inbox (its own
                // included), every CTA polls the local inbox and adds the per-rank sums in rank order
                if (blockIdx.x == 0 && lane < p.size)
                    ll_post(p.peer_inbox[lane] + (size_t)(p.rank * 2 + par) * 2, g, flag);
                double mine = 0.0;
                if (lane < p.size) mine = ll_wait(p.inbox + (size_t)(lane * 2 + par) * 2, flag, p.abort_flag);
                g = 0.0;
                for (int r = 0; r < p.size; ++r) g += __shfl_sync(0xffffffffu, mine, r);
            }
            if (lane == 0) {
                s_coeff = g;
                if (blockIdx.x == 0) p.out[j] = g;
            }
        }
        __syncthreads();
        coeff = s_coeff;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        p.out[p.nq + 1] = (double)*reinterpret_cast<volatile unsigned int *>(p.abort_flag);
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int ensure_sweep_state(pmc_ctx *ctx) {
    if (ctx->sweep_box) return PMC_OK;
    PMC_CHECK_CUDA(cudaSetDevice(ctx->device));
    const size_t box_bytes = (size_t)2 * ctx->sm_count * 2 * sizeof(uint64_t);
    void *mem = nullptr;
    PMC_CHECK_CUDA(cudaMalloc(&mem, box_bytes + 256));
    PMC_CHECK_CUDA(cudaMemset(mem, 0, box_bytes + 256));
    ctx->sweep_box = static_cast<uint64_t *>(mem);
    ctx->sweep_abort = reinterpret_cast<unsigned int *>(static_cast<char *>(mem) + box_bytes);
    return PMC_OK;
}

}  // namespace

extern "C" int pmc_mgs_sweep(pmc_ctx *ctx, const double *Q, int64_t csQ, const double *Z, int64_t csZ, int32_t nq,
                             double *v, int64_t n, const double *w, double *out, uint32_t flags) {
    PMC_REQUIRE(ctx && nq >= 0 && n >= 0 && out, "bad arguments");
    PMC_REQUIRE(n == 0 || v, "NULL vector");
    PMC_REQUIRE(nq == 0 || n == 0 || Q, "NULL panel");
    PMC_REQUIRE(aligned16(v) && aligned16(Q) && aligned16(w) && (csQ % 2) == 0,
                "pmc_mgs_sweep needs 16-byte aligned vectors and an even column stride");
    PMC_REQUIRE(csQ > 0 || nq <= 1, "pmc_mgs_sweep needs a positive column stride");
    PMC_REQUIRE(!Z || (w && aligned16(Z) && (csZ % 2) == 0 && (csZ > 0 || nq <= 1)),
                "a pre-weighted panel needs the weight vector, 16-byte alignment and an even positive stride");
    int rc = ensure_sweep_state(ctx);
    if (rc) return rc;
    SweepParams p;
    memset(&p, 0, sizeof(p));
    p.Q = Q; p.csQ = csQ; p.Z = Z; p.csZ = csZ; p.nq = nq; p.v = v; p.w = w; p.n = n; p.out = out;
    p.cta_box = ctx->sweep_box;
    p.abort_flag = ctx->sweep_abort;
    p.seq_base = ctx->sweep_seq;
    ctx->sweep_seq += (uint32_t)nq + 1u;           // wraps after 2^32 steps; stale flags are long overwritten
    // which GPUs take part is the CALLER's statement (PMC_LOCAL_ONLY: the array lives on this GPU alone, also when
    // the context has peers connected for some other, sharded array)
    const bool shared = ctx->peer_size > 1 && !(flags & PMC_LOCAL_ONLY);
    p.rank = shared ? ctx->peer_rank : 0;
    p.size = shared ? ctx->peer_size : 1;
    // a wait that timed out in an earlier sweep must not abort this one (the flag is reported per call)
    PMC_CHECK_CUDA(cudaMemsetAsync(ctx->sweep_abort, 0, sizeof(unsigned int), ctx->stream));
    p.inbox = ctx->peer_inbox_local;
    for (int r = 0; r < PMC_MAX_PEERS; ++r) p.peer_inbox[r] = ctx->peer_inbox[r];
    p.keep_v = (ctx->l2_keep && n * 8 <= S_KEEP_BYTES) ? 1 : 0;
    p.keep_w = (ctx->l2_keep && w && n * 16 <= S_KEEP_BYTES) ? 1 : 0;
    void *args[] = {&p};
    const dim3 grid(ctx->sm_count), block(S_TPB);
    // cooperative launch: all CTAs are guaranteed to be co-resident (they wait for each other)
    const void *kernel = !w ? (const void *)mgs_sweep_kernel<0>
                         : Z ? (const void *)mgs_sweep_kernel<2> : (const void *)mgs_sweep_kernel<1>;
    PMC_CHECK_CUDA(cudaLaunchCooperativeKernel(kernel, grid, block, args, 0, ctx->stream));
    return pmc_after_launch(ctx, 1, flags);
}

// ---------------------------------------------------------------------------------------------------
// peer inboxes (one process per GPU on one NVSwitch box)
// ---------------------------------------------------------------------------------------------------
extern "C" int pmc_peer_create(pmc_ctx *ctx, int rank, int size, void *ipc_handle_out) {
    PMC_REQUIRE(ctx && ipc_handle_out && size >= 1 && size <= PMC_MAX_PEERS && rank >= 0 && rank < size,
                "bad arguments (at most PMC_MAX_PEERS ranks)");
    PMC_REQUIRE(ctx->peer_inbox_local == nullptr, "peer inbox exists already");
    static_assert(sizeof(cudaIpcMemHandle_t) == PMC_IPC_HANDLE_BYTES, "handle size");
    PMC_CHECK_CUDA(cudaSetDevice(ctx->device));
    void *mem = nullptr;
    const size_t bytes = (size_t)PMC_MAX_PEERS * 2 * 2 * sizeof(uint64_t);
    PMC_CHECK_CUDA(cudaMalloc(&mem, bytes));
    PMC_CHECK_CUDA(cudaMemset(mem, 0, bytes));
    PMC_CHECK_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, mem);
    if (e != cudaSuccess) {
        cudaFree(mem);
        pmc_set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        return PMC_ERR_CUDA;
    }
    memcpy(ipc_handle_out, &h, sizeof(h));
    ctx->peer_inbox_local = static_cast<uint64_t *>(mem);
    ctx->peer_rank = rank;
    ctx->peer_size = 0;                            // becomes `size` in pmc_peer_connect
    ctx->peer_pending_size = size;
    return PMC_OK;
}

extern "C" int pmc_peer_connect(pmc_ctx *ctx, const void *handles) {
    PMC_REQUIRE(ctx && handles && ctx->peer_inbox_local && ctx->peer_pending_size >= 1, "call pmc_peer_create first");
    PMC_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int size = ctx->peer_pending_size;
    for (int r = 0; r < size; ++r) {
        if (r == ctx->peer_rank) {
            ctx->peer_inbox[r] = ctx->peer_inbox_local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)r * sizeof(h), sizeof(h));
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            pmc_set_error("cudaIpcOpenMemHandle (rank %d): %s", r, cudaGetErrorString(e));
            return PMC_ERR_CUDA;
        }
        ctx->peer_inbox[r] = static_cast<uint64_t *>(ptr);
    }
    ctx->peer_size = size;
    // Start the step numbering afresh on every rank: ranks that ran different numbers of (local) sweeps before
    // connecting would otherwise post mismatching flags forever.  Packets of those earlier sweeps are wiped.
    PMC_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->sweep_box)
        PMC_CHECK_CUDA(cudaMemset(ctx->sweep_box, 0, (size_t)2 * ctx->sm_count * 2 * sizeof(uint64_t)));
    ctx->sweep_seq = 0;
    return PMC_OK;
}

extern "C" int pmc_peer_destroy(pmc_ctx *ctx) {
    if (!ctx || !ctx->peer_inbox_local) return PMC_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < PMC_MAX_PEERS; ++r) {
        if (ctx->peer_inbox[r] && ctx->peer_inbox[r] != ctx->peer_inbox_local) cudaIpcCloseMemHandle(ctx->peer_inbox[r]);
        ctx->peer_inbox[r] = nullptr;
    }
    cudaFree(ctx->peer_inbox_local);
    ctx->peer_inbox_local = nullptr;
    ctx->peer_size = ctx->peer_pending_size = 0;
    return PMC_OK;
}
