// gram2_dmma_kernel -- second generation of the tall-skinny Gram / SYRK tile kernel (gram.cu holds the
// first one and the generic kernels).  Same tiling and pipeline (128x128 partial tile per CTA in
// registers, [128 col x 16 row] 128B-swizzled TMA boxes, 6-stage mbarrier ring, mma.sync m8n8k4.f64 =
// DMMA.8x8x4), selected with bit 16 of ctx->gram_opts (env PMC_GRAM_OPTS / pmc_ctx_set_option) until it
// has been measured against gram_dmma_kernel on a B200.  What it changes, all on the FP64 tensor pipe
// that bounds this kernel (round-1 numbers: 128-tile quantisation and diagonal tiles cost 10 %, the
// diagonal weight 6-7 %):
//
//  * diagonal SYRK tiles: the four diagonal 32x32 blocks compute only the 10 of 16 sub-blocks on/below
//    the diagonal, two off-diagonal blocks are K-split over two warps each, so every SMSP carries
//    2.125 block units (a full tile: 4; gram_dmma_kernel: 3);
//  * partially filled blocks (k not a multiple of 32) run only their 8-row / 8-column sub-blocks that
//    hold data, through compile-time specialisations of the stage loop (no predicated DMMA), and are
//    spread over the four SMSPs;
//  * tiles with fewer than 16 blocks of data -- narrow operands such as a panel of block Gram-Schmidt or a
//    few right-hand sides, and the last tile row of a ragged k -- split the K range of every block 4 or 2
//    ways over the otherwise idle warps (each part writes its own slot of the partial tile, the reduction
//    adds them): a 128 x 32 tile keeps all 16 warps and all four tensor pipes busy instead of 4;
//  * the diagonal weight is applied ONCE per stage in shared memory (each warp scales 1/16 of the B tile
//    three stages ahead: 4 DMUL per warp and stage) instead of on every warp's fragments (16 DMUL per
//    warp and stage).  For diagonal tiles the scaled copy goes to the B slot the tile does not load.
//    Products round exactly as in gram_dmma_kernel (fl(w*b) first), so blocks whose K range is not split are
//    bitwise equal (bit 64 of gram_opts disables all K-splits).
//
// The index maps live in gram2_plan.h and are checked on the CPU by tests/native/gram2_model.cpp.
#include "common.cuh"
#include "gram2_plan.h"
#include "gram_params.cuh"

namespace {

constexpr int BM = G2_BM, BN = G2_BN, BK = G2_BK;
constexpr int STAGES = 6;
constexpr int LOOKAHEAD = 3;                      // weight scaling runs this many stages ahead of the MMA
constexpr int THREADS = 32 * G2_WARPS;
constexpr int TILE_ELEMS = G2_TILE_ELEMS;
constexpr uint32_t TILE_BYTES = TILE_ELEMS * 8;
constexpr size_t GRAM2_SMEM = 1024 /*align slack*/ + (size_t)STAGES * (2 * TILE_BYTES + BK * 8) +
                              3 * STAGES * sizeof(uint64_t) + 64;

// One stage (16 rows = KS quarter-steps of 4) of a warp's block.  NM x NN sub-blocks of 8x8; TRI keeps
// the sub-blocks with ni <= mi.  `fo` holds the fragment offsets of the warp's quarter-steps.
// WREG != 0: the diagonal weight of the fragment's row (w_s[wo[ks]]) multiplies the B fragments in registers
// (fl(w * b) first, exactly as gram_dmma_kernel rounds).  WREG == 2 additionally loads the stage's weights up front
// and issues the DMMAs column-of-sub-blocks by column (ni-major): the first NM of them need only b[0], so the DMULs
// of b[1..] -- which queue behind DMMAs on the shared FP64 pipe -- are off the critical path.
template <int NM, int NN, bool TRI, int KS, int WREG>
__device__ __forceinline__ void stage_mma(double (&acc)[4][4][2], const double *a_s, const double *b_s,
                                          const int (&fo)[4], const double *w_s, const int (&wo)[4]) {
    double wv[KS];
    if (WREG == 2) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) wv[ks] = w_s[wo[ks]];
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        double a[NM], b[NN];
#pragma unroll
        for (int ni = 0; ni < NN; ++ni) b[ni] = b_s[ni * 8 * BK + fo[ks]];
#pragma unroll
        for (int mi = 0; mi < NM; ++mi) a[mi] = a_s[mi * 8 * BK + fo[ks]];
        if (WREG == 1) wv[ks] = w_s[wo[ks]];
        if (WREG) {
#pragma unroll
            for (int ni = 0; ni < NN; ++ni) b[ni] *= wv[ks];
        }
        if (WREG == 2) {
#pragma unroll
            for (int ni = 0; ni < NN; ++ni)
#pragma unroll
                for (int mi = 0; mi < NM; ++mi)
                    if (!TRI || ni <= mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        } else {
#pragma unroll
            for (int mi = 0; mi < NM; ++mi)
#pragma unroll
                for (int ni = 0; ni < NN; ++ni)
                    if (!TRI || ni <= mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
}

// Full general tiles with a register-scaled weight: 64 x 16 warp tiles (warp = (wm2, wn2) of a 2 x 8 grid) instead of
// 32 x 32.  Same 16 DMMA per quarter-step and the same 32 accumulators, but only TWO B fragments to scale (2 DMUL
// instead of 4: the weight costs half the FP64-pipe time and half the DMUL -> DMMA dependencies) for two more fragment
// loads.  Sub-block (m8, ni), m8 = 4 h + mi, lives in acc[mi][2 h + ni]; every accumulator receives the same
// products in the same order as in the 32 x 32 layout, so the result is bitwise the same.
__device__ __forceinline__ void stage_mma_wide(double (&acc)[4][4][2], const double *a_s, const double *b_s,
                                               const int (&fo)[4], const double *w_s, const int (&wo)[4]) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        double b[2];
        b[0] = b_s[fo[ks]];
        b[1] = b_s[8 * BK + fo[ks]];
        const double wv = w_s[wo[ks]];
        b[0] *= wv;
        b[1] *= wv;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double a[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = a_s[(h * 4 + mi) * 8 * BK + fo[ks]];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni)
                    dmma884(acc[mi][h * 2 + ni][0], acc[mi][h * 2 + ni][1], a[mi], b[ni]);
        }
    }
}

// shape codes of the warp-uniform dispatch in the stage loop: rectangular blocks (nm-1)*4 + (nn-1), + 16 when the
// warp contracts half of the stage (2 quarter-steps), + 32 for a quarter; triangular (diagonal) blocks 48 + nm-1
constexpr int SHAPE_IDLE = -1, SHAPE_TRI = 48;

// WMODE: 0 no weight, 1 weight applied in shared memory once per stage by the MMA warps themselves (ready ring),
// 2 / 3 weight applied to the fragments in registers (3: weights of a stage loaded up front, ni-major DMMA order),
// 4 weight applied in shared memory by a DEDICATED warpgroup (warp specialisation): warps 16..19 -- one per SMSP --
// scale the B tile of every stage as soon as it has landed (each element once: 16 DMUL per warp and stage instead
// of the 64 that register scaling costs every SMSP, and none of it on the MMA warps' instruction streams), the 16
// MMA warps wait for `ready` instead of `full` and run the unweighted stage loop.  Registers move from the scaler
// warpgroup to the MMA warpgroups with setmaxnreg (640 threads start at 96; the scaler warpgroup shrinks to 32 and thereby releases exactly the 8192 registers the four MMA warpgroups need to grow to 112).
constexpr int WS_THREADS = THREADS + 128;
constexpr int WS_MMA_REGS = 112, WS_SCALER_REGS = 32;

template <int WMODE>
__global__ void __launch_bounds__(WMODE == 4 ? WS_THREADS : THREADS, 1)
gram2_dmma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmW, const GramParams p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024B alignment: the 128B swizzle pattern is a function of the absolute smem address
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    double *sA = reinterpret_cast<double *>(smem);
    double *sB = sA + STAGES * TILE_ELEMS;
    double *sW = sB + STAGES * TILE_ELEMS;
    uint64_t *full = reinterpret_cast<uint64_t *>(sW + STAGES * BK);
    uint64_t *empty = full + STAGES;
    uint64_t *ready = empty + STAGES;               // WMODE 1: stage has been scaled by all warps
    constexpr bool WEIGHTED = WMODE != 0;           // the weight tile is loaded
    constexpr bool WSMEM = WMODE == 1;
    constexpr bool WSPEC = WMODE == 4;              // dedicated scaler warpgroup
    constexpr int WREG = WMODE == 2 ? 1 : WMODE == 3 ? 2 : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int item = blockIdx.x;
    const int tile = item % p.ntiles, split = item / p.ntiles;
    int ti, tj;
    decode_tile(p, tile, ti, tj);
    // Tiles on the diagonal of a symmetric result need their lower part only (`tri_tile`: the triangular warp plan);
    // in a SYRK (symmetric == 1) the A tile serves both sides (`diag`: one operand is loaded), in the lower-only mode
    // of a general product (symmetric == 2: V^T (M V) with symmetric M) the two operands differ.
    const bool tri_tile = p.symmetric != 0 && ti == tj && !(p.opts & G2_OPT_NO_TRI2);
    const bool diag = p.symmetric == 1 && ti == tj;
    const int64_t r0 = (int64_t)split * p.rows_per_split;
    const int64_t r1 = (r0 + p.rows_per_split < p.n) ? r0 + p.rows_per_split : p.n;
    const int nk = (r1 > r0) ? (int)((r1 - r0 + BK - 1) / BK) : 0;

    const int vm = min(BM, p.kA - ti * BM), vn = min(BN, p.kB - tj * BN);   // valid rows / cols of this tile
    const G2Plan plan = g2_plan(warp < G2_WARPS ? warp : 0, diag || tri_tile, vm, vn, p.opts);
    const int shape = !plan.active ? SHAPE_IDLE
                      : plan.tri   ? SHAPE_TRI + plan.nm - 1
                                   : (plan.nm - 1) * 4 + (plan.nn - 1) + (plan.nks == 4 ? 0 : plan.nks == 2 ? 16 : 32);

    if (tid == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (WEIGHTED) tma_prefetch_desc(&tmW);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], G2_WARPS);
            mbar_init(&ready[s], WSPEC ? 4 : G2_WARPS);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (WSPEC) {
        if (warp >= G2_WARPS) {
            // ---- scaler warpgroup: B tile of stage kt *= w (rows of the stage), written to the B slot -- in place for
            // off-diagonal tiles, from the A slot for diagonal tiles (which load one operand only)
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(WS_SCALER_REGS));
            const int st = tid - THREADS;                       // 0..127
            for (int kt = 0; kt < nk; ++kt) {
                const int s = kt % STAGES;
                mbar_wait(&full[s], (kt / STAGES) & 1);
                const double *src = (diag ? sA : sB) + s * TILE_ELEMS;
                double *dst = sB + s * TILE_ELEMS;
                const double *wv = sW + s * BK;
#pragma unroll 2
                for (int part = 0; part < 8; ++part) {
                    // 16-byte chunk q of the swizzled tile: column c = q / 8, physical chunk q % 8 holds rows 2j, 2j+1
                    // with j = (q % 8) ^ (c % 8)
                    const int q = part * 128 + st;
                    const int j = (q & 7) ^ ((q >> 3) & 7);
                    double2 v = *reinterpret_cast<const double2 *>(src + 2 * q);
                    const double2 ww = *reinterpret_cast<const double2 *>(wv + 2 * j);
                    v.x *= ww.x;
                    v.y *= ww.y;
                    *reinterpret_cast<double2 *>(dst + 2 * q) = v;
                }
                fence_proxy_async();      // these generic-proxy writes precede the TMA refill of the slot
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[s]);
            }
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(WS_MMA_REGS));
    }

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

    // WEIGHTED: B tile of stage kt *= w (rows of the stage), written to the B slot -- in place for
    // off-diagonal tiles, from the A slot for diagonal tiles (which load one operand only).  Every warp
    // scales its 8 tile rows, then arrives on ready[stage].
    int sc_elem[2], sc_w[2];
    g2_scale_slice(warp, lane, 0, sc_elem[0], sc_w[0]);
    g2_scale_slice(warp, lane, 1, sc_elem[1], sc_w[1]);
    auto scale_stage = [&](int kt) {
        const int s = kt % STAGES;
        mbar_wait(&full[s], (kt / STAGES) & 1);
        const double *src = (diag ? sA : sB) + s * TILE_ELEMS;
        double *dst = sB + s * TILE_ELEMS;
        const double *wv = sW + s * BK;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            double2 v = *reinterpret_cast<const double2 *>(src + sc_elem[part]);
            const double2 ww = *reinterpret_cast<const double2 *>(wv + sc_w[part]);
            v.x *= ww.x;
            v.y *= ww.y;
            *reinterpret_cast<double2 *>(dst + sc_elem[part]) = v;
        }
        fence_proxy_async();          // these generic-proxy writes precede the TMA refill of the slot
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);
    };
    if (WSMEM) {
        const int pre = nk < LOOKAHEAD ? nk : LOOKAHEAD;
        for (int kt = 0; kt < pre; ++kt) scale_stage(kt);
    }

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    int fo[4], wo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        fo[i] = g2_frag_offset((plan.ks0 + i) & 3, lane);
        wo[i] = ((plan.ks0 + i) & 3) * 4 + (lane & 3);      // stage row of the fragment element: 4 * ks + t
    }
    const int rho = g2_rho(lane >> 2);
    const int rowA = (plan.wm * 32 + rho) * BK;
    const int rowB = (plan.wn * 32 + rho) * BK;
    const double *b_base = (diag && !WSMEM && !WSPEC) ? sA : sB;
    // full general tile, weight scaled in registers: 64 x 16 warp tiles (CTA-uniform choice)
    const bool wide = WREG != 0 && !diag && !tri_tile && vm == BM && vn == BN && !(p.opts & G2_OPT_NO_WIDE);
    const int wide_rowA = ((warp & 1) * 64 + rho) * BK, wide_rowB = ((warp >> 1) * 16 + rho) * BK;
    int fo0[4], wo0[4];                             // fragment / weight offsets of the full K range (ks0 = 0)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        fo0[i] = g2_frag_offset(i, lane);
        wo0[i] = i * 4 + (lane & 3);
    }

    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % STAGES;
        // producer: refill the stage everybody released one iteration ago
        if (tid == 0 && kt >= 1 && kt - 1 + STAGES < nk) {
            const int sp = (kt - 1) % STAGES;
            mbar_wait(&empty[sp], ((kt - 1) / STAGES) & 1);
            issue(kt - 1 + STAGES);
        }
        mbar_wait((WSMEM || WSPEC) ? &ready[s] : &full[s], (kt / STAGES) & 1);
        const double *a_s = sA + s * TILE_ELEMS + rowA;
        const double *b_s = b_base + s * TILE_ELEMS + rowB;
        const double *w_s = sW + s * BK;
        if (WREG != 0 && wide) {
            stage_mma_wide(acc, sA + s * TILE_ELEMS + wide_rowA, sB + s * TILE_ELEMS + wide_rowB, fo0, w_s, wo0);
        } else
        switch (shape) {
#define G2_RECT(NM, NN)                                                                                   \
    case (NM - 1) * 4 + (NN - 1): stage_mma<NM, NN, false, 4, WREG>(acc, a_s, b_s, fo, w_s, wo); break;                  \
    case (NM - 1) * 4 + (NN - 1) + 16: stage_mma<NM, NN, false, 2, WREG>(acc, a_s, b_s, fo, w_s, wo); break;             \
    case (NM - 1) * 4 + (NN - 1) + 32: stage_mma<NM, NN, false, 1, WREG>(acc, a_s, b_s, fo, w_s, wo); break;
            G2_RECT(4, 4) G2_RECT(4, 3) G2_RECT(4, 2) G2_RECT(4, 1)
            G2_RECT(3, 4) G2_RECT(3, 3) G2_RECT(3, 2) G2_RECT(3, 1)
            G2_RECT(2, 4) G2_RECT(2, 3) G2_RECT(2, 2) G2_RECT(2, 1)
            G2_RECT(1, 4) G2_RECT(1, 3) G2_RECT(1, 2) G2_RECT(1, 1)
#undef G2_RECT
            case SHAPE_TRI + 3: stage_mma<4, 4, true, 4, WREG>(acc, a_s, b_s, fo, w_s, wo); break;
            case SHAPE_TRI + 2: stage_mma<3, 3, true, 4, WREG>(acc, a_s, b_s, fo, w_s, wo); break;
            case SHAPE_TRI + 1: stage_mma<2, 2, true, 4, WREG>(acc, a_s, b_s, fo, w_s, wo); break;
            case SHAPE_TRI + 0: stage_mma<1, 1, true, 4, WREG>(acc, a_s, b_s, fo, w_s, wo); break;
            default: break;
        }
        if (WSMEM && kt + LOOKAHEAD < nk) scale_stage(kt + LOOKAHEAD);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // epilogue: partial tile, column-major [j][i]; sub-blocks nobody reads later are not written
    if (WREG != 0 && wide) {
        double *out = p.partials + (size_t)item * (BM * BN);
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int h = q >> 1, ni = q & 1;
                const int i = (warp & 1) * 64 + (h * 4 + mi) * 8 + g2_rho(g);
                const int j = (warp >> 1) * 16 + ni * 8;
                out[(size_t)(j + g2_rho(2 * t)) * BM + i] = acc[mi][q][0];
                out[(size_t)(j + g2_rho(2 * t + 1)) * BM + i] = acc[mi][q][1];
            }
    } else if (plan.active) {
        double *out = p.partials + (size_t)item * (BM * BN);
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                if (mi >= plan.nm || ni >= plan.nn || (plan.tri && ni > mi)) continue;
                out[g2_out_index(plan, mi, ni, lane, 0)] = acc[mi][ni][0];
                out[g2_out_index(plan, mi, ni, lane, 1)] = acc[mi][ni][1];
            }
    }
}

// stage 2: out[i + j*ldo] = sum over the splits of (sum over the K parts of the entry's block), mirrored for SYRK.
// Fixed order => bitwise reproducible.
__global__ void gram2_reduce_kernel(const double *__restrict__ partials, int nsplit, GramParams p,
                                    double *__restrict__ out, int64_t ldo, int accumulate) {
    const int tile = blockIdx.y;
    int ti, tj;
    decode_tile(p, tile, ti, tj);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= BM * BN) return;
    const int il = e % BM, jl = e / BM;
    const int gi = ti * BM + il, gj = tj * BN + jl;
    if (gi >= p.kA || gj >= p.kB) return;
    // (same rule as in the tile kernel: the triangular plan -- and its slot layout -- on every diagonal tile of a
    // symmetric result)
    const bool diag = p.symmetric != 0 && ti == tj && (p.symmetric == 1 || !(p.opts & G2_OPT_NO_TRI2));
    if (p.symmetric && ti == tj && gi < gj) return; // strictly-upper part of diagonal tiles: mirrored
    const size_t tile_elems = (size_t)BM * BN;
    const double *src = partials + (size_t)tile * tile_elems;
    const size_t step = (size_t)p.ntiles * tile_elems;
    const int vm = min(BM, p.kA - ti * BM), vn = min(BN, p.kB - tj * BN);
    // Where the entry's K parts were stored (up to 4 partial-tile entries): diagonal tiles keep a block in place
    // (+ the transposed slot for the second half of a split block); other tiles number their (block, K part)
    // units and store unit u in slot u.
    int src_idx[4], parts = 1;
    src_idx[0] = e;
    if (diag) {
        if (g2_is_split_block(il >> 5, jl >> 5, vm, p.opts)) {
            parts = 2;
            src_idx[1] = g2_mirror_index(il, jl);
        }
    } else {
        int first;
        g2_general_sources(il, jl, vm, vn, p.opts, first, parts);
#pragma unroll
        for (int q = 0; q < 4; ++q) src_idx[q] = g2_slot_index(first + (q < parts ? q : 0), il, jl);
    }
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) {
        const double *t = src + (size_t)sp * step;
        double v = t[src_idx[0]];
        if (parts > 1) v += t[src_idx[1]];
        if (parts > 2) v = (v + t[src_idx[2]]) + t[src_idx[3]];
        s += v;
    }
    if (accumulate) s += out[gi + (int64_t)gj * ldo];
    out[gi + (int64_t)gj * ldo] = s;
    if (p.symmetric && gi != gj) out[gj + (int64_t)gi * ldo] = s;
}

}  // namespace

// Launch of the tile kernel + reduction on behalf of pmc_gram_impl (which has validated the operands,
// chosen the split and encoded the tensor maps).  Two launches.
int pmc_gram2_launch(pmc_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmW,
                     const GramParams &p, int64_t items, int64_t nsplit, bool weighted, double *out,
                     int64_t ldo, int accumulate) {
    const int tslot = pmc_timing_begin(ctx, PMC_KERNEL_GRAM);
    if (weighted && (p.opts & G2_OPT_W_SMEM)) {
        PMC_CHECK_CUDA(cudaFuncSetAttribute(gram2_dmma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)GRAM2_SMEM));
        gram2_dmma_kernel<1><<<(unsigned)items, THREADS, GRAM2_SMEM, ctx->stream>>>(tmA, tmB, tmW, p);
    } else if (weighted && (p.opts & G2_OPT_W_WARPSPEC)) {
        PMC_CHECK_CUDA(cudaFuncSetAttribute(gram2_dmma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)GRAM2_SMEM));
        gram2_dmma_kernel<4><<<(unsigned)items, WS_THREADS, GRAM2_SMEM, ctx->stream>>>(tmA, tmB, tmW, p);
    } else if (weighted && (p.opts & G2_OPT_W_NIMAJOR)) {
        PMC_CHECK_CUDA(cudaFuncSetAttribute(gram2_dmma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)GRAM2_SMEM));
        gram2_dmma_kernel<3><<<(unsigned)items, THREADS, GRAM2_SMEM, ctx->stream>>>(tmA, tmB, tmW, p);
    } else if (weighted) {
        PMC_CHECK_CUDA(cudaFuncSetAttribute(gram2_dmma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)GRAM2_SMEM));
        gram2_dmma_kernel<2><<<(unsigned)items, THREADS, GRAM2_SMEM, ctx->stream>>>(tmA, tmB, tmW, p);
    } else {
        PMC_CHECK_CUDA(cudaFuncSetAttribute(gram2_dmma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)GRAM2_SMEM));
        gram2_dmma_kernel<0><<<(unsigned)items, THREADS, GRAM2_SMEM, ctx->stream>>>(tmA, tmB, tmW, p);
    }
    pmc_timing_end(ctx, tslot);
    PMC_CHECK_CUDA(cudaGetLastError());
    dim3 rgrid((BM * BN + 255) / 256, p.ntiles);
    gram2_reduce_kernel<<<rgrid, 256, 0, ctx->stream>>>(p.partials, (int)nsplit, p, out, ldo, accumulate);
    return PMC_OK;
}
