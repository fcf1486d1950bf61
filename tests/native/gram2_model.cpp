// Host-side model of gram2_dmma_kernel + gram2_reduce_kernel (pymor_b200/csrc/gram2.cu) -- TEST
// INFRASTRUCTURE ONLY, compiled with g++ by tests/test_gram2_model.py.
//
// The build container has no GPU.  This model executes, lane by lane, what the kernel's threads do with
// the index maps of pymor_b200/csrc/gram2_plan.h (the very header the kernel includes): TMA boxes land
// in 128B-swizzled shared-memory tiles, every warp scales its slice of the stage by the weights, loads
// its DMMA fragments, the m8n8k4 MMA combines them (a = A[g][t], b = B[t][g], c = C[g][2t], C[g][2t+1]
// for lane 4g+t), the epilogue scatters the accumulators into the partial tile, the reduction sums the
// splits and the mirrored K-halves.  The result must equal A^T diag(w) B.  What it cannot check is the
// mbarrier protocol (no concurrency here); partial tiles start as NaN so that any read of an entry no
// warp wrote shows up.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../../pymor_b200/csrc/gram2_plan.h"

namespace {

struct Params {
    int64_t n, rows_per_split;
    int tiles_m, tiles_n, ntiles, symmetric, kA, kB, opts;
};

// same numbering as decode_tile (gram_params.cuh)
void decode_tile(const Params &p, int tile, int &ti, int &tj) {
    if (p.symmetric) {
        int rem = tile;
        tj = 0;
        while (rem >= p.tiles_m - tj) { rem -= p.tiles_m - tj; ++tj; }
        ti = tj + rem;
    } else {
        ti = tile % p.tiles_m;
        tj = tile / p.tiles_m;
    }
}

// cp.async.bulk.tensor 2D box [128 cols x 16 rows] at (row d0, column c0), 128B swizzle, OOB -> 0
void tma_box(double *tile, const double *P, int64_t cs, int k, int64_t n, int64_t d0, int c0) {
    for (int c = 0; c < G2_BM; ++c)
        for (int kk = 0; kk < G2_BK; ++kk) {
            const int64_t d = d0 + kk;
            const int col = c0 + c;
            tile[g2_tile_offset(c, kk)] = (d < n && col < k) ? P[(int64_t)col * cs + d] : 0.0;
        }
}

}  // namespace

// dmma_per_smsp: optional [ntiles][4] DMMA instruction counts per SMSP (warp % 4) for one stage of each tile
extern "C" int g2_model_gram(const double *A, int64_t csA, int kA, const double *B, int64_t csB, int kB, int64_t n,
                             const double *w, double *out, int64_t ldo, int symmetric, int opts,
                             int64_t rows_per_split, int64_t *dmma_per_smsp) {
    Params p;
    p.n = n; p.rows_per_split = rows_per_split; p.symmetric = symmetric; p.kA = kA; p.kB = kB; p.opts = opts;
    p.tiles_m = (kA + G2_BM - 1) / G2_BM;
    p.tiles_n = (kB + G2_BN - 1) / G2_BN;
    p.ntiles = symmetric ? p.tiles_m * (p.tiles_m + 1) / 2 : p.tiles_m * p.tiles_n;
    if (rows_per_split % G2_BK) return -1;
    const int64_t nsplit = (n + rows_per_split - 1) / rows_per_split;
    const int64_t items = nsplit * p.ntiles;
    const size_t tile_elems = (size_t)G2_BM * G2_BN;
    std::vector<double> partials((size_t)items * tile_elems, std::numeric_limits<double>::quiet_NaN());
    const bool weighted = w != nullptr;

    std::vector<double> sA(G2_TILE_ELEMS), sB(G2_TILE_ELEMS), sW(G2_BK);
    for (int64_t item = 0; item < items; ++item) {
        const int tile = (int)(item % p.ntiles);
        const int64_t split = item / p.ntiles;
        int ti, tj;
        decode_tile(p, tile, ti, tj);
        const bool diag = symmetric && ti == tj;
        const int64_t r0 = split * rows_per_split;
        const int64_t r1 = (r0 + rows_per_split < n) ? r0 + rows_per_split : n;
        const int nk = (r1 > r0) ? (int)((r1 - r0 + G2_BK - 1) / G2_BK) : 0;
        const int vm = g2_min(G2_BM, kA - ti * G2_BM), vn = g2_min(G2_BN, kB - tj * G2_BN);

        // per-lane accumulators of the 16 warps
        std::vector<double> acc((size_t)G2_WARPS * 32 * 4 * 4 * 2, 0.0);
        auto ACC = [&](int warp, int lane, int mi, int ni, int c) -> double & {
            return acc[((((size_t)warp * 32 + lane) * 4 + mi) * 4 + ni) * 2 + c];
        };
        G2Plan plans[G2_WARPS];
        for (int warp = 0; warp < G2_WARPS; ++warp) plans[warp] = g2_plan(warp, diag, vm, vn, opts);

        for (int kt = 0; kt < nk; ++kt) {
            const int64_t d0 = r0 + (int64_t)kt * G2_BK;
            // issue(): A box, B box unless diagonal, 16 weights
            tma_box(sA.data(), A, csA, kA, n, d0, ti * G2_BM);
            if (!diag) tma_box(sB.data(), B, csB, kB, n, d0, tj * G2_BN);
            else std::fill(sB.begin(), sB.end(), std::numeric_limits<double>::quiet_NaN());   // slot not loaded
            if (weighted)
                for (int kk = 0; kk < G2_BK; ++kk) sW[kk] = (d0 + kk < n) ? w[d0 + kk] : 0.0;
            // scale_stage(): every warp, every lane, two 16-byte chunks
            if (weighted) {
                const std::vector<double> src = diag ? sA : sB;
                for (int warp = 0; warp < G2_WARPS; ++warp)
                    for (int lane = 0; lane < 32; ++lane)
                        for (int part = 0; part < 2; ++part) {
                            int elem, wofs;
                            g2_scale_slice(warp, lane, part, elem, wofs);
                            sB[elem] = src[elem] * sW[wofs];
                            sB[elem + 1] = src[elem + 1] * sW[wofs + 1];
                        }
            }
            const double *b_base = (diag && !weighted) ? sA.data() : sB.data();
            for (int warp = 0; warp < G2_WARPS; ++warp) {
                const G2Plan &pl = plans[warp];
                if (!pl.active) continue;
                for (int q = 0; q < pl.nks; ++q) {
                    const int ks = (pl.ks0 + q) & 3;
                    double a[32][4], b[32][4];
                    for (int lane = 0; lane < 32; ++lane) {
                        const int rho = g2_rho(lane >> 2);
                        const int rowA = (pl.wm * 32 + rho) * G2_BK, rowB = (pl.wn * 32 + rho) * G2_BK;
                        const int fo = g2_frag_offset(ks, lane);
                        for (int mi = 0; mi < pl.nm; ++mi) a[lane][mi] = sA[rowA + mi * 8 * G2_BK + fo];
                        for (int ni = 0; ni < pl.nn; ++ni) b[lane][ni] = b_base[rowB + ni * 8 * G2_BK + fo];
                    }
                    for (int mi = 0; mi < pl.nm; ++mi)
                        for (int ni = 0; ni < pl.nn; ++ni) {
                            if (pl.tri && ni > mi) continue;
                            if (dmma_per_smsp && kt == 0 && split == 0) dmma_per_smsp[tile * 4 + (warp & 3)] += 1;
                            // mma.sync.m8n8k4: C[g][2t+c] += sum_t' A[g][t'] * B[t'][2t+c]
                            for (int g = 0; g < 8; ++g)
                                for (int col = 0; col < 8; ++col) {
                                    double s = ACC(warp, 4 * g + (col >> 1), mi, ni, col & 1);
                                    for (int t = 0; t < 4; ++t) s = std::fma(a[4 * g + t][mi], b[4 * col + t][ni], s);
                                    ACC(warp, 4 * g + (col >> 1), mi, ni, col & 1) = s;
                                }
                        }
                }
            }
        }
        // epilogue
        double *po = partials.data() + (size_t)item * tile_elems;
        for (int warp = 0; warp < G2_WARPS; ++warp) {
            const G2Plan &pl = plans[warp];
            if (!pl.active) continue;
            for (int lane = 0; lane < 32; ++lane)
                for (int mi = 0; mi < 4; ++mi)
                    for (int ni = 0; ni < 4; ++ni) {
                        if (mi >= pl.nm || ni >= pl.nn || (pl.tri && ni > mi)) continue;
                        for (int c = 0; c < 2; ++c) {
                            double &dst = po[g2_out_index(pl, mi, ni, lane, c)];
                            if (!std::isnan(dst)) return -2;          // two writers for one partial entry
                            dst = ACC(warp, lane, mi, ni, c);
                        }
                    }
        }
    }

    // gram2_reduce_kernel
    for (int tile = 0; tile < p.ntiles; ++tile) {
        int ti, tj;
        decode_tile(p, tile, ti, tj);
        const bool diag = symmetric && ti == tj;
        const int vm = g2_min(G2_BM, kA - ti * G2_BM), vn = g2_min(G2_BN, kB - tj * G2_BN);
        for (int e = 0; e < G2_BM * G2_BN; ++e) {
            const int il = e % G2_BM, jl = e / G2_BM;
            const int gi = ti * G2_BM + il, gj = tj * G2_BN + jl;
            if (gi >= kA || gj >= kB) continue;
            if (diag && gi < gj) continue;
            int src_idx[4], parts = 1;
            src_idx[0] = e;
            if (diag) {
                if (g2_is_split_block(il >> 5, jl >> 5, vm, opts)) {
                    parts = 2;
                    src_idx[1] = g2_mirror_index(il, jl);
                }
            } else {
                int first;
                g2_general_sources(il, jl, vm, vn, opts, first, parts);
                for (int q = 0; q < 4; ++q) src_idx[q] = g2_slot_index(first + (q < parts ? q : 0), il, jl);
            }
            double s = 0.0;
            for (int64_t sp = 0; sp < nsplit; ++sp) {
                const double *src = partials.data() + ((size_t)sp * p.ntiles + tile) * tile_elems;
                double v = src[src_idx[0]];
                if (parts > 1) v += src[src_idx[1]];
                if (parts > 2) v = (v + src[src_idx[2]]) + src[src_idx[3]];
                s += v;
            }
            out[gi + (int64_t)gj * ldo] = s;
            if (symmetric && gi != gj) out[gj + (int64_t)gi * ldo] = s;
        }
    }
    return 0;
}

// shared-memory address (in doubles, relative to the stage's tile) of a lane's fragment load
extern "C" int g2_model_frag_addr(int wblk, int sub, int ks, int lane) {
    return (wblk * 32 + g2_rho(lane >> 2)) * G2_BK + sub * 8 * G2_BK + g2_frag_offset(ks, lane);
}
// same, through the tile layout function (row, K index): must agree with the line above
extern "C" int g2_model_frag_addr_by_layout(int wblk, int sub, int ks, int lane) {
    return g2_tile_offset(g2_frag_row(wblk, sub, lane), ks * 4 + (lane & 3));
}
extern "C" void g2_model_scale_slice(int warp, int lane, int part, int *elem, int *wofs) {
    g2_scale_slice(warp, lane, part, *elem, *wofs);
}
