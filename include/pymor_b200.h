/*
 * pymor_b200.h -- C ABI of libpymor_b200.so: device-resident fp64 snapshot linear algebra for
 * pyMOR's VectorArray hot path on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  Every entry point replaces one method of the reference's
 * CPU backend `NumpyVectorArrayImpl` (src/pymor/vectorarrays/numpy.py) or one default method of
 * `Operator` (src/pymor/operators/interface.py); the file:line it replaces is cited at each
 * declaration.  The host side (pymor_b200/vectorarray.py, ctypes) binds exactly these symbols.
 *
 * Conventions
 * -----------
 *  - All data are IEEE fp64.  A *panel* is a set of `k` columns of `n` contiguous doubles each:
 *    column c starts at `ptr + c*cs` (cs = column stride in doubles; any sign; 0 = broadcast one
 *    column).  This is the reference's Fortran-ordered (dim, len) ndarray (numpy.py:16-18) with the
 *    view arithmetic (slice start/step) already folded into `ptr`/`cs` by the caller.
 *  - Device memory is allocated and owned by the caller (torch); the library only borrows raw
 *    pointers for the duration of a call.  All work is enqueued on the context's stream.
 *  - `out*` result pointers must be device-accessible: device memory (multi-GPU partials that are
 *    all-reduced afterwards) or pinned/mapped host memory (single-GPU: the kernel writes the
 *    result straight to the host).  With PMC_SYNC the stream is synchronised before returning.
 *  - Return value: 0 on success, negative pmc_status on failure; pmc_last_error() gives the
 *    message (thread-local).  No exceptions cross the boundary, no global state besides pmc_ctx.
 *  - `n` is the LOCAL number of rows (DOFs) of this rank's shard; reductions return local
 *    partial sums (the reference's MPI backend does the same, vectorarrays/mpi.py:394-416).
 */
#ifndef PYMOR_B200_H
#define PYMOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMC_VERSION 100

typedef enum {
    PMC_OK = 0,
    PMC_ERR_INVALID = -1,     /* bad argument */
    PMC_ERR_CUDA = -2,        /* CUDA runtime/driver error (message in pmc_last_error) */
    PMC_ERR_WORKSPACE = -3,   /* workspace too small; pmc_ctx_workspace_needed() tells how much */
    PMC_ERR_UNSUPPORTED = -4
} pmc_status;

/* flags */
#define PMC_SYNC 1u        /* synchronise the stream before returning (results visible on host) */
#define PMC_SYMMETRIC 2u   /* pmc_gram: B is the same panel as A -> SYRK (lower tiles + mirror)  */
#define PMC_FORCE_SIMPLE 4u /* use the generic DFMA kernel even where the TMA/DMMA path applies (tests) */
#define PMC_LOCAL_ONLY 16u /* pmc_mgs_sweep: this array lives on this GPU alone -- no sums with the connected peer GPUs */
#define PMC_LOWER_ONLY 8u  /* pmc_gram / pmc_csr_gram: the caller knows the (square) result is symmetric although the two
                              operands differ (V^T (M V) with symmetric M): contract only the tiles on/below the diagonal
                              and mirror -- half the flop */

typedef struct pmc_ctx pmc_ctx;

int pmc_version(void);
const char *pmc_last_error(void);

/* Context: one per (device, stream).  `stream` is a cudaStream_t (0 = legacy default stream). */
int pmc_ctx_create(int device, void *stream, pmc_ctx **out);
int pmc_ctx_destroy(pmc_ctx *ctx);
/* Caller-owned scratch (device memory) for split partials; must be 256-byte aligned. */
int pmc_ctx_set_workspace(pmc_ctx *ctx, void *dev_ptr, size_t bytes);
size_t pmc_ctx_workspace_needed(pmc_ctx *ctx); /* bytes asked for by the last PMC_ERR_WORKSPACE */
int pmc_ctx_sync(pmc_ctx *ctx);
int pmc_ctx_sm_count(pmc_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t pmc_ctx_launch_count(pmc_ctx *ctx);
/* Tuning switches of the kernels (defaults come from the environment when the context is created):
 * "gram_opts" (bit set of pymor_b200/csrc/gram2_plan.h; 16 selects the second-generation Gram tile kernel),
 * "gram_waves" (target CTA waves of the split-n Gram kernel), "gram_l2_mb" (L2 budget that bounds the rows per split),
 * "csr_variant", "l2_keep", "kernel_timing".  No reference
 * counterpart; results do not depend on them beyond the summation order of the Gram partials. */
int pmc_ctx_set_option(pmc_ctx *ctx, const char *name, int value);
int pmc_ctx_get_option(pmc_ctx *ctx, const char *name, int *value);
/* Per-launch kernel times.  With the option "kernel_timing" set to 1 the library brackets every launch of its
 * tensor-core kernels (the Gram tile kernel: PMC_KERNEL_GRAM; the block lincomb kernel: PMC_KERNEL_LINCOMB) and of
 * the CSR SpMM (PMC_KERNEL_SPMM) with CUDA events on the launching stream.  This call waits for the recorded
 * launches, writes (kind, milliseconds) pairs in launch order, stores their number in *count and clears the log
 * (at most PMC_TIMING_SLOTS = 1024 launches are kept).  No reference counterpart: measurement only (bench.py's
 * `roofline` is the time of the kernel itself, not of the call around it). */
#define PMC_KERNEL_GRAM 0
#define PMC_KERNEL_LINCOMB 1
#define PMC_KERNEL_SPMM 2
int pmc_ctx_kernel_times(pmc_ctx *ctx, int32_t *kinds_out, double *ms_out, int32_t cap, int32_t *count);
/* Pin caller-owned host memory (a shared-memory segment common to the ranks of one box) so that
 * reduction kernels can write their tiny partial results straight into it; *dev_ptr_out is the address
 * to pass as `out`.  The ranks then sum the partials on the host -- the replacement for the
 * Gather-to-root + host sum of vectorarrays/mpi.py:394-416 when the payload is a few doubles. */
int pmc_host_register(pmc_ctx *ctx, void *host_ptr, size_t bytes, void **dev_ptr_out);
int pmc_host_unregister(pmc_ctx *ctx, void *host_ptr);

/* inner / gramian / apply2 with a diagonal product.
 * out[i + j*ldo] = sum_d A[d + i*csA] * w[d] * B[d + j*csB]   (w == NULL: Euclidean)
 * Replaces NumpyVectorArrayImpl.inner (vectorarrays/numpy.py:155-163), VectorArrayImpl.gramian
 * (vectorarrays/interface.py:1068-1069) and Operator.apply2 for a diagonal matrix operator
 * (operators/interface.py:79-109 with operators/numpy.py:232-237).
 * Large, TMA-describable panels run on the TMA + FP64 tensor-core (DMMA) kernel; everything else
 * on the generic kernel.  PMC_SYMMETRIC computes only the lower tile triangle and mirrors, so the
 * result is bitwise symmetric like the reference's dsyrk path. */
int pmc_gram(pmc_ctx *ctx, const double *A, int64_t csA, int32_t kA, const double *B, int64_t csB,
             int32_t kB, int64_t n, const double *w, double *out, int64_t ldo, uint32_t flags);

/* lincomb:  R[d + j*csR] = sum_i A[d + i*csA] * CT[i + j*ldc],  j < m.
 * CT is DEVICE memory holding the (k x m) coefficient matrix in Fortran order (column j of the
 * coefficients contiguous), ldc >= k.  Replaces NumpyVectorArrayImpl.lincomb
 * (vectorarrays/numpy.py:175-180). */
int pmc_lincomb(pmc_ctx *ctx, const double *A, int64_t csA, int32_t k, int64_t n, const double *CT,
                int64_t ldc, int32_t m, double *R, int64_t csR, uint32_t flags);

/* pairwise_inner / pairwise_apply2 with a diagonal product:
 * out[c] = sum_d A[d + c*csA] * w[d] * B[d + c*csB].
 * Replaces NumpyVectorArrayImpl.pairwise_inner (vectorarrays/numpy.py:165-173) and
 * Operator.pairwise_apply2 (operators/interface.py:111-143). */
int pmc_pairwise_inner(pmc_ctx *ctx, const double *A, int64_t csA, const double *B, int64_t csB,
                       int32_t k, int64_t n, const double *w, double *out, uint32_t flags);

/* norm2: out[c] = sum_d w[d] * A[d + c*csA]^2.  Replaces NumpyVectorArrayImpl.norm2/norm
 * (vectorarrays/numpy.py:182-194). */
int pmc_norm2(pmc_ctx *ctx, const double *A, int64_t csA, int32_t k, int64_t n, const double *w,
              double *out, uint32_t flags);

/* axpy: Y[:, c] += alpha[c or 0] * X[:, c]  (csX == 0 broadcasts one x column).
 * alpha is a HOST array of n_alpha (1 or k) doubles.  Replaces NumpyVectorArrayImpl.axpy
 * (vectorarrays/numpy.py:112-138). */
int pmc_axpy(pmc_ctx *ctx, double *Y, int64_t csY, const double *X, int64_t csX, int32_t k,
             int64_t n, const double *alpha, int32_t n_alpha, uint32_t flags);

/* Fused Gram-Schmidt step on single vectors: Y += alpha[0]*X, then out[0] = sum_d Q[d]*w[d]*Y[d]
 * (Q == NULL: out[0] = sum_d w[d]*Y[d]^2) in one pass.  Bitwise identical to pmc_axpy followed by
 * pmc_pairwise_inner(Q, Y) / pmc_norm2(Y); the host layer uses it to merge `v.axpy(-p, q_j)` with the
 * following `q_{j+1}.pairwise_inner(v)` / `v.norm()` of gram_schmidt.py:84-91. */
int pmc_axpy_dot(pmc_ctx *ctx, double *Y, const double *X, const double *alpha, const double *Q,
                 int64_t n, const double *w, double *out, uint32_t flags);

/* One whole Gram-Schmidt sweep in one launch (persistent cooperative kernel, csrc/mgs.cu):
 *   for j in 0..nq-1:  r[j] = sum_d Q[d + j*csQ] * w[d] * v[d]  (summed over all ranks);  v -= r[j] * Q[:, j]
 *   norm2 = sum_d w[d] * v[d]^2  (summed over all ranks)
 * out (nq + 2 doubles, device-accessible, e.g. pinned host memory) receives r[0..nq-1], norm2 and a status
 * word (0 = ok, 1 = a wait inside the kernel timed out).  Replaces the inner loop of gram_schmidt
 * (algorithms/gram_schmidt.py:84-93: pairwise_inner + axpy per pair, then norm) and, on row-sharded arrays,
 * the Gather + host sum per pair of vectorarrays/mpi.py:394-416: after pmc_peer_connect the cross-GPU sum of
 * every coefficient happens inside the kernel over NVLink peer memory; all ranks must call it with the same
 * nq in the same order and get bitwise identical r / norm2.  Vectors must be 16-byte aligned, csQ even.
 * Z (optional, needs w): panel of the pre-weighted columns Z[:, j] = w * Q[:, j] (pmc_diag_apply, once per
 * column when it becomes final); the dots then read Z[:, j] instead of Q[:, j] and w, which takes the weight
 * stream out of every step (r[j] = sum_d Z[d + j*csZ] * v[d]). */
int pmc_mgs_sweep(pmc_ctx *ctx, const double *Q, int64_t csQ, const double *Z, int64_t csZ, int32_t nq,
                  double *v, int64_t n, const double *w, double *out, uint32_t flags);

/* Peer inboxes for the in-kernel cross-GPU sums of pmc_mgs_sweep (one process per GPU, one box).
 * pmc_peer_create allocates this rank's inbox and returns its CUDA IPC handle (PMC_IPC_HANDLE_BYTES bytes);
 * the caller exchanges the handles of all ranks (any transport; pymor_b200 uses torch.distributed's
 * all_gather_object) and passes them, concatenated in rank order, to pmc_peer_connect. */
#define PMC_MAX_PEERS 8
#define PMC_IPC_HANDLE_BYTES 64
int pmc_peer_create(pmc_ctx *ctx, int rank, int size, void *ipc_handle_out);
int pmc_peer_connect(pmc_ctx *ctx, const void *handles);
int pmc_peer_destroy(pmc_ctx *ctx);

/* scal: Y[:, c] *= alpha[c or 0].  Replaces NumpyVectorArrayImpl.scal (numpy.py:85-100). */
int pmc_scal(pmc_ctx *ctx, double *Y, int64_t csY, int32_t k, int64_t n, const double *alpha,
             int32_t n_alpha, uint32_t flags);

/* copy of a strided panel: D[:, c] = S[:, c].  Replaces the slicing copies in
 * NumpyVectorArrayImpl.copy/append (numpy.py:58-83). */
int pmc_copy(pmc_ctx *ctx, double *D, int64_t csD, const double *S, int64_t csS, int32_t k,
             int64_t n, uint32_t flags);

/* gather of arbitrary (possibly repeated) columns: D[:, c] = base[:, idx[c]] with leading dimension
 * ld; idx is a HOST array.  Replaces fancy indexing `self._array[:, ind]` (numpy.py:21, :59). */
int pmc_gather_cols(pmc_ctx *ctx, double *D, int64_t csD, const double *base, int64_t ld,
                    const int64_t *idx, int32_t k, int64_t n, uint32_t flags);

/* In-place compaction: columns [first, first + count) of the panel move `shift` <= first columns down
 * (A[:, c - shift] = A[:, c] for increasing c; ranges may overlap).  What `del array[...]` needs after the removed
 * vectors are gone (NumpyVectorArrayImpl.delete re-allocates instead, vectorarrays/numpy.py:47-57): one launch per
 * run of kept vectors. */
int pmc_shift_cols(pmc_ctx *ctx, double *A, int64_t cs, int64_t first, int64_t count, int64_t shift,
                   int64_t n, uint32_t flags);

/* dofs: out[i + c*ndofs] = A[dof[i] + c*csA]; dof is a HOST array of local row indices.
 * Replaces NumpyVectorArrayImpl.dofs (numpy.py:196-201). */
int pmc_dofs(pmc_ctx *ctx, const double *A, int64_t csA, int32_t k, int64_t n, const int64_t *dof,
             int32_t ndofs, double *out, uint32_t flags);

/* amax: per column the FIRST row index of maximal |value| and that value.
 * Replaces NumpyVectorArrayImpl.amax (numpy.py:203-211). */
int pmc_amax(pmc_ctx *ctx, const double *A, int64_t csA, int32_t k, int64_t n, int64_t *out_idx,
             double *out_val, uint32_t flags);

/* Y[:, c] = w .* X[:, c]: Operator.apply of a diagonal matrix operator
 * (operators/numpy.py:232-237 for a scipy dia/csr diagonal matrix). */
int pmc_diag_apply(pmc_ctx *ctx, const double *w, const double *X, int64_t csX, double *Y,
                   int64_t csY, int32_t k, int64_t n, uint32_t flags);

/* CSR operator.apply:  Y[r + c*csY] = sum_{p in row r} vals[p] * x(colidx[p], c).
 * rowptr (int64, nrows+1), colidx (int32), vals (fp64) are DEVICE arrays of this rank's row block.
 * Column indices < nrows address the rank's own rows, x = X[col + c*csX]; indices >= nrows address
 * halo rows owned by other ranks, x = ghost[(col - nrows)*k + c] (row-major receive buffer filled by
 * the caller's neighbour exchange; NULL when the block has no halo).  Replaces
 * NumpyMatrixOperator.apply for scipy sparse matrices (operators/numpy.py:232-237) and the
 * "MPI-aware operator" the reference delegates to external solvers (operators/mpi.py:22-26). */
int pmc_csr_apply(pmc_ctx *ctx, int64_t nrows, const int64_t *rowptr, const int32_t *colidx,
                  const double *vals, const double *X, int64_t csX, const double *ghost, double *Y,
                  int64_t csY, int32_t k, uint32_t flags);

/* Halo packing: out[r*k + c] = A[idx[r] + c*csA] for a DEVICE index array (the rows a neighbour
 * rank asked for), row-major so that one contiguous message per neighbour results. */
int pmc_gather_rows(pmc_ctx *ctx, const double *A, int64_t csA, int32_t k, int64_t n,
                    const int64_t *idx_dev, int64_t nidx, double *out, uint32_t flags);

/* Fused CSR two-form  out[i + j*ldo] = sum_r V[r + i*csV] * (M U)[r, j]  without materialising the
 * n x k product M U: row panels of M U are formed in the workspace and contracted at once.
 * `out` must be DEVICE memory (it is accumulated panel by panel); `ghost` as in pmc_csr_apply (for U).
 * Replaces Operator.apply2 = V.inner(op.apply(U)) (operators/interface.py:79-109). */
int pmc_csr_gram(pmc_ctx *ctx, int64_t nrows, const int64_t *rowptr, const int32_t *colidx,
                 const double *vals, const double *V, int64_t csV, int32_t kV, const double *U,
                 int64_t csU, int32_t kU, const double *ghost, double *out, int64_t ldo,
                 uint32_t flags);

/* Fused CSR pairwise two-form  out[j] = sum_r V[r + j*csV] * (M U)[r, j],  j < k  (Operator.pairwise_apply2 =
 * V.pairwise_inner(op.apply(U)), operators/interface.py:111-143 -- e.g. the product norms ||u_j||_M^2 that
 * VectorArray.norm2(product) asks for): one pass over M, U and V, the n x k product M U is never written.
 * `out`: k doubles, device or pinned host memory; block partials are summed in a fixed order (bitwise
 * reproducible).  `ghost` as in pmc_csr_apply (for U). */
int pmc_csr_pairwise(pmc_ctx *ctx, int64_t nrows, const int64_t *rowptr, const int32_t *colidx,
                     const double *vals, const double *V, int64_t csV, const double *U, int64_t csU,
                     const double *ghost, int32_t k, double *out, uint32_t flags);

/* Collectives of the row-sharded arrays (one process per GPU on one box).  The reference gathers the per-rank
 * partial results on rank 0 and sums them on the host (vectorarrays/mpi.py:394-416, MPI Gather of pickled arrays);
 * here every rank holds its k1 x k2 (or k) partial in device memory and pmc_allreduce_sum adds them in place over
 * NVLink (ncclAllReduce, fp64, on the context's stream, i.e. ordered behind the reduction kernel that wrote `buf`).
 * NCCL is looked up in the process at run time (no link-time dependency).  pmc_comm_unique_id: 128 bytes created by
 * one rank, to be handed to all ranks by the host; pmc_comm_init is collective over the `size` ranks. */
#define PMC_COMM_ID_BYTES 128
int pmc_comm_unique_id(void *id_out);
int pmc_comm_init(pmc_ctx *ctx, int rank, int size, const void *id_bytes);
int pmc_comm_destroy(pmc_ctx *ctx);
int pmc_comm_size(pmc_ctx *ctx); /* 0 without a communicator */
int pmc_allreduce_sum(pmc_ctx *ctx, double *buf, int64_t count, uint32_t flags);
/* G[j + i*ld] = G[i + j*ld] for i > j (device memory, column-major k x k): restores the bitwise symmetry of a
 * Gramian after the cross-GPU sum, whose chunks are added in different rank orders (the reference's syrk result and
 * its rank-ordered host sum are symmetric: vectorarrays/numpy.py:155-163, mpi.py:394-416). */
int pmc_mirror_lower(pmc_ctx *ctx, double *G, int64_t k, int64_t ld, uint32_t flags);

/* Roofline denominators measured on the spot: sustained FP64 tensor-core (DMMA) and DFMA
 * throughput in TFLOP/s over `ms` milliseconds of back-to-back issue on every SM.  use_dmma: 0 = DFMA, 1 = DMMA,
 * 2 = half of the warps DMMA / half DFMA, 3 = every warp interleaves both (2 and 3 report the combined rate: do the
 * two instruction kinds share the FP64 units?). */
int pmc_measure_fp64_peak(pmc_ctx *ctx, int use_dmma, double ms, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* PYMOR_B200_H */
