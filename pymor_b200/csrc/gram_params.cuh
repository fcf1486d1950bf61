// Launch parameters and tile numbering shared by the Gram kernels (gram.cu, gram2.cu).
#pragma once
#include "common.cuh"

struct GramParams {
    int64_t n;
    int64_t rows_per_split;   // multiple of BK
    int tiles_m, tiles_n, ntiles;
    int symmetric;
    int kA, kB;
    int opts;                 // switches (env PMC_GRAM_OPTS): 1 no diag skip, 2 no empty-block skip, 4 no transpose; >= 16: gram2_plan.h
    double *partials;         // [item][BN][BM]
};

static __device__ __forceinline__ void decode_tile(const GramParams &p, int tile, int &ti, int &tj) {
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

// gram2.cu: launch of gram2_dmma_kernel + gram2_reduce_kernel (two launches) for pmc_gram_impl
int pmc_gram2_launch(pmc_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmW,
                     const GramParams &p, int64_t items, int64_t nsplit, bool weighted, double *out,
                     int64_t ldo, int accumulate);
