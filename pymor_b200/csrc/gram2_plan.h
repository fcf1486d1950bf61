// Index maps of gram2_dmma_kernel (gram2.cu), shared between the kernel and the host-side model
// that checks them without a GPU (tests/native/gram2_model.cpp).  Everything here is pure integer
// arithmetic: which warp computes which 32x32 block of a 128x128 Gram tile (and which 8-row
// sub-blocks / K quarter-steps of it), where a fragment element sits in the 128B-swizzled
// shared-memory tile, where an accumulator goes in the partial tile, and which slice of a stage a
// warp scales by the diagonal weight.
#pragma once

#if defined(__CUDACC__)
#define G2_HD __host__ __device__ __forceinline__
#else
#define G2_HD inline
#endif

constexpr int G2_BM = 128, G2_BN = 128, G2_BK = 16;
constexpr int G2_WARPS = 16;
constexpr int G2_TILE_ELEMS = G2_BM * G2_BK;   // one operand tile of one stage: [128 cols][16 rows]

// switches (ctx->gram_opts, env PMC_GRAM_OPTS); bits 1, 2, 4 belong to gram_dmma_kernel (gram.cu)
constexpr int G2_OPT_ENABLE = 16;      // use gram2_dmma_kernel instead of gram_dmma_kernel
constexpr int G2_OPT_NO_TRI = 32;      // diagonal blocks of diagonal SYRK tiles: all 16 sub-blocks
constexpr int G2_OPT_NO_KSPLIT = 64;   // no K-split of two off-diagonal blocks of diagonal tiles
constexpr int G2_OPT_NO_RAGGED = 128;  // partially filled blocks run in full
constexpr int G2_OPT_NO_TRI2 = 4096;    // lower-only general product: diagonal tiles run in full instead of on the triangular plan (A/B)
constexpr int G2_OPT_NO_WIDE = 2048;    // register-scaled weight: no 64x16 warp tiles on full general tiles (A/B)
constexpr int G2_OPT_W_WARPSPEC = 1024; // weight applied in shared memory by a dedicated scaler warpgroup (warp specialisation, setmaxnreg)
constexpr int G2_OPT_W_NIMAJOR = 512;  // register-scaled weight: stage weights loaded up front, ni-major DMMA order (A/B)
constexpr int G2_OPT_W_SMEM = 256;     // diagonal weight applied once per stage in shared memory (third barrier ring) instead of
                                       // on each warp's B fragments in registers; measured slower on B200 (profiles/r02_summary.md)

// What one warp does in one Gram tile.
//   (wm, wn)  32x32 block of the tile: rows wm*32.. (columns of A), columns wn*32.. (columns of B)
//   nm, nn    8-row / 8-column sub-blocks of it that hold data (1..4)
//   ks0, nks  quarter-steps (4 rows of the 16-row stage each) this warp contracts: [ks0, ks0+nks)
//   tri       diagonal block of a diagonal SYRK tile: only sub-blocks (mi, ni) with ni <= mi
//   sm, sn    32x32 slot of the partial tile that receives the accumulators.  A block whose K range is
//             split over several warps occupies several slots (the reduction adds them): slots of blocks
//             the tile does not have -- the upper triangle of a diagonal tile, the empty block rows /
//             columns of a narrow tile.
struct G2Plan {
    int wm, wn, nm, nn, ks0, nks, tri, sm, sn, active;
};

G2_HD int g2_min(int a, int b) { return a < b ? a : b; }

// 8-wide sub-blocks with data in block `w` (0..3) of a tile side holding `valid` rows
G2_HD int g2_sub_blocks(int valid, int w, int opts) {
    const int rem = valid - w * 32;
    if (rem <= 0) return 0;
    if (opts & G2_OPT_NO_RAGGED) return 4;
    return g2_min(4, (rem + 7) >> 3);
}

// b-th strictly-lower block of the 4x4 block triangle, row by row: (1,0) (2,0) (2,1) (3,0) (3,1) (3,2)
G2_HD void g2_lower_block(int b, int &wm, int &wn) {
    wm = (b >= 3) ? 3 : (b >= 1) ? 2 : 1;
    wn = b - wm * (wm - 1) / 2;
}

// The two blocks of a diagonal tile whose K range is split over two warps: (3,1) and (3,2)
G2_HD bool g2_is_split_block(int wm, int wn, int vm, int opts) {
    if (opts & G2_OPT_NO_KSPLIT) return false;
    if (wm != 3 || (wn != 1 && wn != 2)) return false;
    return g2_sub_blocks(vm, wm, opts) == 4 && g2_sub_blocks(vm, wn, opts) == 4;
}

// Off-diagonal / general tiles: bm x bn blocks hold data (narrow operands -- a panel of a block
// Gram-Schmidt, a few right-hand sides, the last tile row of a ragged k -- have fewer than 16).  Then the K
// range of every block is split f = 4 or 2 ways so that (almost) all 16 warps and all four FP64 tensor
// pipes work: unit u = block * f + part runs on warp u and writes slot u of the partial tile.
// Blocks are numbered so that the blocks of a SHORT block row (ragged M) sit on four different SMSPs (n_fast: the
// block column runs fastest); a short block column (ragged N) is spread by the default numbering.
G2_HD void g2_tile_shape(int vm, int vn, int opts, int &bm, int &bn, int &f, int &n_fast) {
    bm = g2_min(4, (vm + 31) >> 5);
    bn = g2_min(4, (vn + 31) >> 5);
    f = 1;
    if (!(opts & G2_OPT_NO_KSPLIT)) {
        if (bm * bn * 4 <= G2_WARPS) f = 4;
        else if (bm * bn * 2 <= G2_WARPS) f = 2;
    }
    n_fast = (vm & 31) != 0 ? 1 : 0;
}
G2_HD int g2_block_number(int wm, int wn, int bm, int bn, int n_fast) { return n_fast ? wn + bn * wm : wm + bm * wn; }

// Diagonal SYRK tiles need the 10 blocks on/below the diagonal.  Each SMSP (warp % 4) has its own
// FP64 tensor pipe, so the work is dealt out per SMSP: one diagonal block (10 of 16 sub-blocks),
// one full off-diagonal block and one K-half of a split off-diagonal block -- 2.125 block units
// per SMSP instead of 4 for a full tile.  Other tiles: see g2_tile_shape; consecutive units sit on
// different SMSPs, so short (ragged) blocks and K parts spread evenly.
G2_HD G2Plan g2_plan(int warp, bool diag, int vm, int vn, int opts) {
    G2Plan p;
    p.ks0 = 0; p.nks = 4; p.tri = 0; p.active = 1;
    if (diag) {
        int mirror = 0;
        if (warp < 4) {
            p.wm = p.wn = warp;
            p.tri = (opts & G2_OPT_NO_TRI) ? 0 : 1;
        } else if (warp < 8) {
            g2_lower_block(warp - 4, p.wm, p.wn);
        } else if (warp < 12) {
            g2_lower_block(4 + ((warp - 8) >> 1), p.wm, p.wn);
            const int half = (warp - 8) & 1;
            if (g2_is_split_block(p.wm, p.wn, vm, opts)) {
                p.ks0 = 2 * half; p.nks = 2; mirror = half;
            } else if (half) {
                p.active = 0;
            }
        } else {
            p.wm = p.wn = 0;
            p.active = 0;
        }
        p.sm = mirror ? p.wn : p.wm;          // second K-half: the transposed slot, unused by a diagonal tile
        p.sn = mirror ? p.wm : p.wn;
    } else {
        int bm, bn, f, n_fast;
        g2_tile_shape(vm, vn, opts, bm, bn, f, n_fast);
        const int block = warp / f, part = warp % f;
        if (block >= bm * bn) {
            p.wm = p.wn = 0;
            p.active = 0;
        } else {
            p.wm = n_fast ? block / bn : block % bm;
            p.wn = n_fast ? block % bn : block / bm;
            p.nks = 4 / f;
            p.ks0 = part * p.nks;
        }
        p.sm = warp & 3;
        p.sn = warp >> 2;
    }
    p.nm = g2_sub_blocks(vm, p.wm, opts);
    p.nn = g2_sub_blocks(vn, p.wn, opts);
    if (p.nm == 0 || p.nn == 0) p.active = 0;
    return p;
}

// Fragment row g (0..7) of an 8-row sub-block is tile row rho(g): with the 128B TMA swizzle the 16
// lanes of each half-warp then hit 16 distinct 8-byte bank pairs.
G2_HD int g2_rho(int g) { return 2 * (g & 3) + (g >> 2); }

// A tile row holds the 16 stage rows (K) of one column: 128 bytes = eight 16-byte chunks, chunk q
// stored at position q ^ (row & 7) (CU_TENSOR_MAP_SWIZZLE_128B, tile base 1024B-aligned).
G2_HD int g2_tile_offset(int row, int kk /*0..15*/) {
    return row * G2_BK + ((((kk >> 1) ^ (row & 7)) << 1) | (kk & 1));
}

// Offset (doubles, relative to the sub-block's first row rho) of lane's fragment element for
// quarter-step ks: K index inside the stage is ks*4 + t, t = lane & 3.
G2_HD int g2_frag_offset(int ks, int lane) {
    const int rho = g2_rho(lane >> 2);
    return ((((ks * 2 + ((lane >> 1) & 1)) ^ rho) << 1) + (lane & 1));
}
G2_HD int g2_frag_row(int wblk, int sub, int lane) { return wblk * 32 + sub * 8 + g2_rho(lane >> 2); }

// Accumulator c (0/1) of sub-block (mi, ni) of lane: element of the 128x128 partial tile, stored
// column-major [j][i], inside the warp's slot.
G2_HD int g2_out_index(const G2Plan &p, int mi, int ni, int lane, int c) {
    const int i = p.sm * 32 + mi * 8 + g2_rho(lane >> 2);
    const int j = p.sn * 32 + ni * 8 + g2_rho(2 * (lane & 3) + c);
    return j * G2_BM + i;
}

// Where the second K-half of element (il, jl) of a diagonal tile was stored (the transposed slot)
G2_HD int g2_mirror_index(int il, int jl) {
    const int i2 = (jl >> 5) * 32 + (il & 31), j2 = (il >> 5) * 32 + (jl & 31);
    return j2 * G2_BM + i2;
}

// Off-diagonal / general tile: element (il, jl) of the tile is the sum of the entries g2_slot_index(slot, il, jl)
// of the slots first .. first+count-1 (one per K part of its block).
G2_HD void g2_general_sources(int il, int jl, int vm, int vn, int opts, int &first, int &count) {
    int bm, bn, f, n_fast;
    g2_tile_shape(vm, vn, opts, bm, bn, f, n_fast);
    first = g2_block_number(il >> 5, jl >> 5, bm, bn, n_fast) * f;   // units first .. first+f-1 = slots of the same number
    count = f;
}
G2_HD int g2_slot_index(int slot, int il, int jl) {
    return ((slot >> 2) * 32 + (jl & 31)) * G2_BM + (slot & 3) * 32 + (il & 31);
}

// Diagonal weight: warp `warp` scales tile rows 8*warp .. 8*warp+7 of a stage, two 16-byte chunks
// per lane (part = 0, 1).  elem = offset of the chunk in the tile, wofs = offset of its two weights
// in the stage's 16 weights.
G2_HD void g2_scale_slice(int warp, int lane, int part, int &elem, int &wofs) {
    const int row = 8 * warp + (lane >> 3) + 4 * part;
    const int pos = lane & 7;
    elem = row * G2_BK + 2 * pos;
    wofs = 2 * (pos ^ (row & 7));
}
