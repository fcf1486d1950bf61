// Context, error reporting and TMA descriptor encoding for libpymor_b200.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void pmc_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int pmc_version(void) { return PMC_VERSION; }
extern "C" const char *pmc_last_error(void) { return g_err; }

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int pmc_ctx_create(int device, void *stream, pmc_ctx **out) {
    PMC_REQUIRE(out != nullptr, "out is NULL");
    int ndev = 0;
    PMC_CHECK_CUDA(cudaGetDeviceCount(&ndev));
    PMC_REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
    PMC_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PMC_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        pmc_set_error("libpymor_b200 is built for sm_100a only; device %d is sm_%d%d", device,
                      prop.major, prop.minor);
        return PMC_ERR_UNSUPPORTED;
    }
    pmc_ctx *c = new pmc_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->stream = reinterpret_cast<cudaStream_t>(stream);
    c->sm_count = prop.multiProcessorCount;
    c->n_tickets = 4096;
    c->partials_cap = 1 << 16;
    if (cudaMalloc(&c->tickets, c->n_tickets * sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(c->tickets, 0, c->n_tickets * sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(&c->partials, (size_t)c->partials_cap * 16) != cudaSuccess) {
        pmc_set_error("cannot allocate ticket counters / block partials");
        delete c;
        return PMC_ERR_CUDA;
    }
    cudaDriverEntryPointQueryResult q;
    void *fn = nullptr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr) {
        pmc_set_error("cuTensorMapEncodeTiled not available from the driver");
        cudaFree(c->tickets);
        cudaFree(c->partials);
        delete c;
        return PMC_ERR_CUDA;
    }
    c->encode_tiled = fn;
    const char *nk = getenv("PMC_NO_L2_KEEP");
    c->l2_keep = (nk && nk[0] == '1') ? 0 : 1;
    const char *go = getenv("PMC_GRAM_OPTS");
    c->gram_opts = go ? atoi(go) : 16;   // 16 = G2_OPT_ENABLE: gram2_dmma_kernel (A/B on B200: profiles/r02_summary.md); 0 = gram_dmma_kernel
    const char *cv = getenv("PMC_CSR_VARIANT");
    c->csr_variant = cv ? atoi(cv) : 0;
    const char *gw = getenv("PMC_GRAM_WAVES");
    c->gram_waves = (gw && atoi(gw) > 0) ? atoi(gw) : 32;
    const char *gl = getenv("PMC_GRAM_L2_MB");
    c->gram_l2_mb = gl ? atoi(gl) : 48;
    *out = c;
    return PMC_OK;
}

extern "C" int pmc_ctx_destroy(pmc_ctx *ctx) {
    if (!ctx) return PMC_OK;
    cudaSetDevice(ctx->device);
    pmc_peer_destroy(ctx);
    pmc_comm_destroy(ctx);
    for (int i = 0; i < PMC_TIMING_SLOTS; ++i)
        for (int j = 0; j < 2; ++j)
            if (ctx->t_ev[i][j]) cudaEventDestroy(ctx->t_ev[i][j]);
    cudaFree(ctx->sweep_box);
    cudaFree(ctx->tickets);
    cudaFree(ctx->partials);
    delete ctx;
    return PMC_OK;
}

extern "C" int pmc_ctx_set_workspace(pmc_ctx *ctx, void *dev_ptr, size_t bytes) {
    PMC_REQUIRE(ctx != nullptr, "ctx is NULL");
    PMC_REQUIRE((reinterpret_cast<uintptr_t>(dev_ptr) & 255) == 0, "workspace must be 256B aligned");
    ctx->ws = dev_ptr;
    ctx->ws_bytes = bytes;
    return PMC_OK;
}

extern "C" size_t pmc_ctx_workspace_needed(pmc_ctx *ctx) { return ctx ? ctx->ws_needed : 0; }

extern "C" int pmc_ctx_sync(pmc_ctx *ctx) {
    PMC_REQUIRE(ctx != nullptr, "ctx is NULL");
    PMC_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PMC_OK;
}

// Tuning switches, initialised from the environment at context creation (PMC_GRAM_OPTS, PMC_GRAM_WAVES,
// PMC_CSR_VARIANT, PMC_NO_L2_KEEP); settable at run time so that one process can A/B kernel variants.
static int *pmc_option_slot(pmc_ctx *ctx, const char *name) {
    if (!ctx || !name) return nullptr;
    if (!strcmp(name, "gram_opts")) return &ctx->gram_opts;
    if (!strcmp(name, "gram_waves")) return &ctx->gram_waves;
    if (!strcmp(name, "gram_l2_mb")) return &ctx->gram_l2_mb;
    if (!strcmp(name, "csr_variant")) return &ctx->csr_variant;
    if (!strcmp(name, "l2_keep")) return &ctx->l2_keep;
    if (!strcmp(name, "kernel_timing")) return &ctx->kernel_timing;
    return nullptr;
}

extern "C" int pmc_ctx_set_option(pmc_ctx *ctx, const char *name, int value) {
    int *slot = pmc_option_slot(ctx, name);
    PMC_REQUIRE(slot != nullptr, "unknown option (gram_opts, gram_waves, gram_l2_mb, csr_variant, l2_keep, kernel_timing)");
    PMC_REQUIRE(slot != &ctx->gram_waves || value > 0, "gram_waves must be positive");
    *slot = value;
    return PMC_OK;
}

extern "C" int pmc_ctx_get_option(pmc_ctx *ctx, const char *name, int *value) {
    int *slot = pmc_option_slot(ctx, name);
    PMC_REQUIRE(slot != nullptr && value != nullptr, "unknown option (gram_opts, gram_waves, gram_l2_mb, csr_variant, l2_keep, kernel_timing)");
    *value = *slot;
    return PMC_OK;
}

int pmc_timing_begin(pmc_ctx *ctx, int kind) {
    if (!ctx->kernel_timing || ctx->t_count >= PMC_TIMING_SLOTS) return -1;
    const int slot = ctx->t_count;
    for (int j = 0; j < 2; ++j)
        if (!ctx->t_ev[slot][j] && cudaEventCreate(&ctx->t_ev[slot][j]) != cudaSuccess) return -1;
    if (cudaEventRecord(ctx->t_ev[slot][0], ctx->stream) != cudaSuccess) return -1;
    ctx->t_kind[slot] = kind;
    ctx->t_count = slot + 1;
    return slot;
}

void pmc_timing_end(pmc_ctx *ctx, int slot) {
    if (slot >= 0) cudaEventRecord(ctx->t_ev[slot][1], ctx->stream);
}

extern "C" int pmc_ctx_kernel_times(pmc_ctx *ctx, int32_t *kinds_out, double *ms_out, int32_t cap,
                                    int32_t *count) {
    PMC_REQUIRE(ctx && count && (cap == 0 || (kinds_out && ms_out)), "bad arguments");
    const int n = ctx->t_count < cap ? ctx->t_count : cap;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        PMC_CHECK_CUDA(cudaEventSynchronize(ctx->t_ev[i][1]));
        PMC_CHECK_CUDA(cudaEventElapsedTime(&ms, ctx->t_ev[i][0], ctx->t_ev[i][1]));
        kinds_out[i] = ctx->t_kind[i];
        ms_out[i] = (double)ms;
    }
    *count = n;
    ctx->t_count = 0;
    return PMC_OK;
}

extern "C" int pmc_ctx_sm_count(pmc_ctx *ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" int64_t pmc_ctx_launch_count(pmc_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pmc_encode_tmap_f64(pmc_ctx *ctx, CUtensorMap *tm, const void *base, int rank, uint64_t dim0,
                        uint64_t dim1, uint64_t stride1_bytes, uint32_t box0, uint32_t box1,
                        bool swizzle128) {
    cuuint64_t gdim[2] = {dim0, dim1};
    cuuint64_t gstr[1] = {stride1_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    encode_tiled_fn fn = reinterpret_cast<encode_tiled_fn>(ctx->encode_tiled);
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<void *>(base),
                    gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        pmc_set_error("cuTensorMapEncodeTiled failed (CUresult %d): base=%p rank=%d dims=(%llu,%llu) "
                      "stride=%llu box=(%u,%u)",
                      (int)r, base, rank, (unsigned long long)dim0, (unsigned long long)dim1,
                      (unsigned long long)stride1_bytes, box0, box1);
        return PMC_ERR_CUDA;
    }
    return PMC_OK;
}

// Pin caller-owned host memory (e.g. a POSIX shared-memory segment shared by the ranks of one box) and
// return the address kernels can write to.  Used for the host-side sum of tiny per-rank partials.
extern "C" int pmc_host_register(pmc_ctx *ctx, void *host_ptr, size_t bytes, void **dev_ptr_out) {
    PMC_REQUIRE(ctx && host_ptr && bytes > 0 && dev_ptr_out, "bad arguments");
    PMC_CHECK_CUDA(cudaSetDevice(ctx->device));
    PMC_CHECK_CUDA(cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    void *d = nullptr;
    PMC_CHECK_CUDA(cudaHostGetDevicePointer(&d, host_ptr, 0));
    *dev_ptr_out = d;
    return PMC_OK;
}

extern "C" int pmc_host_unregister(pmc_ctx *ctx, void *host_ptr) {
    PMC_REQUIRE(ctx && host_ptr, "bad arguments");
    PMC_CHECK_CUDA(cudaHostUnregister(host_ptr));
    return PMC_OK;
}
