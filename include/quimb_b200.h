/*
 * quimb_b200 -- C ABI of the B200-native tensor-network contraction engine.
 *
 * This header is the drop-in boundary.  Every entry point replaces one
 * library call that quimb's hot path bottoms out in (all reference citations
 * are relative to the quimb source tree):
 *
 *   qb_contract_pair      <- autoray do("tensordot") / do("einsum") issued by the
 *                            cotengra pairwise loop under
 *                            quimb/tensor/contraction.py:272-292 (array_contract)
 *                            and by Tensor.__matmul__, quimb/tensor/tensor_core.py:3786-3808
 *   qb_permute            <- do("transpose")+do("reshape") materialisation in
 *                            quimb/tensor/array_ops.py:148-180 (fuse)
 *   qb_svd_trunc          <- svd_truncated incl. the keep rule, renormalisation
 *                            and absorb, quimb/tensor/decomp.py:829-1118, :901-1029,
 *                            :693-721 (qb_svd + qb_svals_to_keep are its parts)
 *   qb_qr_stab            <- qr_stabilized, quimb/tensor/decomp.py:2055-2216
 *   qb_dot / qb_axpby / qb_scale / qb_multi_dot / qb_multi_axpy
 *                         <- ARPACK dsaupd vector algebra behind
 *                            quimb/linalg/scipy_linalg.py:113-128 (eigs_scipy);
 *                            the Krylov control loop itself is host Python
 *                            (quimb_b200/lanczos.py) -- there is no qb_lanczos_*
 *                            and no plan_create/execute/destroy entry: cached
 *                            chains are CUDA graphs captured by the host layer
 *
 * Conventions
 *   - plain C, no C++/torch types; device pointers are borrowed, never freed
 *     or retained by the library beyond the call.
 *   - all strides are in ELEMENTS (not bytes) and may describe any
 *     non-overlapping strided view (torch.Tensor.stride()).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - return value: 0 = success; <0 = invalid argument (-(index+1));
 *     >0 = CUDA / numerical failure.  qb_last_error() gives a message.
 *   - thread safety: no global mutable state except the thread-local error
 *     string; calls on different streams may run concurrently.
 */
#ifndef QUIMB_B200_H
#define QUIMB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB_MAX_RANK 32
#define QB_ABI_VERSION 1

typedef enum {
  QB_F32 = 0,
  QB_F64 = 1,
  QB_C64 = 2,
  QB_C128 = 3
} qb_dtype_t;

/* A strided device tensor view (mirrors what quimb keeps in Tensor._data). */
typedef struct {
  void *ptr;
  int32_t dtype; /* qb_dtype_t */
  int32_t rank;
  int64_t shape[QB_MAX_RANK];
  int64_t stride[QB_MAX_RANK]; /* elements */
} qb_tensor_t;

/* quimb's absorb codes, decomp.py:201-211 (QB_ABSORB_FULL stands for None). */
enum {
  QB_ABSORB_FULL = 100, /* 'U,s,VH' : return the three parts */
  QB_ABSORB_S = 2,      /* 's' */
  QB_ABSORB_LSQRT = -12,
  QB_ABSORB_RORTHOG = -11,
  QB_ABSORB_LFACTOR = -10,
  QB_ABSORB_LEFT = -1,
  QB_ABSORB_BOTH = 0,
  QB_ABSORB_RIGHT = 1,
  QB_ABSORB_LORTHOG = 10,
  QB_ABSORB_RFACTOR = 11,
  QB_ABSORB_RSQRT = 12
};

/* quimb's cutoff modes, decomp.py:265-270 */
enum {
  QB_CUTOFF_ABS = 1,
  QB_CUTOFF_REL = 2,
  QB_CUTOFF_SUM2 = 3,
  QB_CUTOFF_RSUM2 = 4,
  QB_CUTOFF_SUM1 = 5,
  QB_CUTOFF_RSUM1 = 6
};

/* contraction engine selection for qb_contract_pair */
enum {
  QB_ENGINE_AUTO = 0, /* heuristic                                    */
  QB_ENGINE_DMMA = 1, /* native fp64 tensor-core path (DMMA)          */
  QB_ENGINE_OZAKI = 2, /* tcgen05 int8 error-free-split path (fp64)    */
  QB_ENGINE_STREAM = 3, /* HBM-bound streaming path for small-operator
                         * steps (N <= 16 and K <= 16 after mode grouping:
                         * gate application); other shapes take the DMMA
                         * path.  Opt-in until timed on a B200.          */
  /* OR-able flag: the caller guarantees that bytes [0, 1024) of `workspace`
   * (the stream-K flag words) were zero before its first use and are written
   * by this library only -- which always leaves them zero again -- so the
   * per-launch clear of those words is skipped (saves one memset node per
   * contraction; matters for back-to-back launches and CUDA graphs). */
  QB_ENGINE_WS_ZEROED = 0x100
};

/* ---- library ---------------------------------------------------------- */
int qb_abi_version(void);
const char *qb_last_error(void);
/* number of kernels launched by this library in this process (monotonic) */
int64_t qb_launch_count(void);

/* ---- pairwise contraction --------------------------------------------- */
/*
 * C[labelsC] = sum over labels not in C of  op(A)[labelsA] * op(B)[labelsB]
 *
 * Labels are arbitrary int32 mode names (the integer image of quimb's index
 * names).  A label in A and B but not C is contracted; in A, B and C it is a
 * batch ("hyper") index; in only one input and not C it is summed over; a
 * label repeated inside one input takes that input's diagonal.  Every label
 * of C must appear in A or B.  No operand is transposed or copied: the index
 * permutation is folded into the kernel's tile loads / stores.
 * conjA / conjB conjugate complex inputs on load.
 * workspace may be NULL when qb_contract_pair_workspace() returns 0.
 */
int qb_contract_pair(const qb_tensor_t *A, const int32_t *labelsA,
                     const qb_tensor_t *B, const int32_t *labelsB,
                     qb_tensor_t *C, const int32_t *labelsC, int conjA,
                     int conjB, int engine, void *workspace,
                     size_t workspace_bytes, void *stream);

/* as qb_contract_pair, but C = alpha * op(A).op(B) + beta * C (real scalars;
 * beta != 0 accumulates into the existing contents of C).  This is the
 * building block of the Lanczos orthogonalisation and of the blocked
 * Householder updates. */
int qb_contract_pair_ab(const qb_tensor_t *A, const int32_t *labelsA,
                        const qb_tensor_t *B, const int32_t *labelsB,
                        qb_tensor_t *C, const int32_t *labelsC, int conjA,
                        int conjB, double alpha, double beta, void *workspace,
                        size_t workspace_bytes, void *stream);

int64_t qb_contract_pair_workspace(const qb_tensor_t *A,
                                   const int32_t *labelsA,
                                   const qb_tensor_t *B,
                                   const int32_t *labelsB, const qb_tensor_t *C,
                                   const int32_t *labelsC, int engine);

/*
 * Host-only: run the planner and report the GEMM view it derived.
 * out[0..7] = M, N, K, batch, n_m_modes, n_n_modes, n_k_modes, n_batch_modes
 * out[8] = tile config id, out[9] = split-k factor, out[10] = vecA,
 * out[11] = vecB, out[12] = vecC.  Needs no GPU; used by the CPU test-suite.
 */
int qb_contract_pair_plan(const qb_tensor_t *A, const int32_t *labelsA,
                          const qb_tensor_t *B, const int32_t *labelsB,
                          const qb_tensor_t *C, const int32_t *labelsC,
                          int64_t *out16);

/*
 * TEST / debug entry, host pointers only: runs the streaming engine's row
 * code (shared __host__ __device__ functions of contract_stream.cu) in a
 * host loop so that its index bookkeeping can be checked without a device.
 * Never called by the product; returns -9 for shapes the engine does not take.
 */
int qb_debug_contract_stream_host(const qb_tensor_t *A, const int32_t *labelsA,
                                  const qb_tensor_t *B, const int32_t *labelsB,
                                  qb_tensor_t *C, const int32_t *labelsC,
                                  int conjA, int conjB, double alpha,
                                  double beta);

/*
 * Many independent same-signature contractions in one launch: A[i], B[i], C[i]
 * share shape/stride/labels (taken from A0/B0/C0) but have their own base
 * pointers given in device arrays of `count` pointers.
 */
int qb_contract_batched(const qb_tensor_t *A0, const int32_t *labelsA,
                        const qb_tensor_t *B0, const int32_t *labelsB,
                        qb_tensor_t *C0, const int32_t *labelsC,
                        const void *const *dA, const void *const *dB,
                        void *const *dC, int64_t count, int conjA, int conjB,
                        void *stream);

/* ---- layout / element-wise -------------------------------------------- */
/* dst (any strides) = src (any strides), same shape; optional conjugation */
int qb_permute(const qb_tensor_t *src, qb_tensor_t *dst, int conj,
               void *stream);
/* y = alpha * x + beta * y over n contiguous elements (alpha,beta: re,im) */
int qb_axpby(int dtype, int64_t n, const double alpha[2], const void *x,
             const double beta[2], void *y, void *stream);
/* x *= alpha / (*dev_scalar)  (dev_scalar may be NULL); Lanczos normalise */
int qb_scale(int dtype, int64_t n, const double alpha[2],
             const void *dev_div_scalar, void *x, void *stream);
/* y = x * alpha / (*dev_div_scalar) (divisor optional; zero divisor -> zeros):
 * the normalised copy of a Krylov residual in one pass. */
int qb_scale_into(int dtype, int64_t n, double alpha, const void *dev_div_scalar,
                  const void *x, void *y, void *stream);
/* out[0] (device, same dtype as x; complex: conj(x).y) = <x|y>; deterministic
 * two-stage reduction; workspace >= qb_dot_workspace(n) bytes */
int qb_dot(int dtype, int64_t n, const void *x, const void *y, void *out,
           void *workspace, void *stream);
int64_t qb_dot_workspace(int64_t n);
/* Krylov block algebra (one pass over V each; f64): V is m x n row-major with
 * leading dimension ldv, m <= 16.
 *   qb_multi_dot :  out[j] = <V[j], w>            (device out, deterministic)
 *   qb_multi_axpy:  w += alpha * sum_j h[j] V[j]  (h on the device)
 * Together: one classical Gram-Schmidt pass of the Lanczos vector w. */
int qb_multi_dot(int dtype, int m, int64_t n, const void *V, int64_t ldv,
                 const void *w, void *out, void *workspace, void *stream);
int64_t qb_multi_dot_workspace(void);
int qb_multi_axpy(int dtype, int m, int64_t n, const void *V, int64_t ldv,
                  const void *h, double alpha, void *w, void *stream);
/* x (rows x cols, row-major contiguous) *= d[col]^p  (side=1) or d[row]^p
 * (side=0); p in {1, 0.5}; d is real (f32 for F32/C64, f64 otherwise) */
int qb_scale_diag(int dtype, int64_t rows, int64_t cols, void *x,
                  const void *d, int side, int sqrt_d, void *stream);

/* contiguous precision conversion f32<->f64, c64<->c128 (n elements).  f32 /
 * c64 operands are widened (exactly) for the fp64 engines and the result is
 * rounded once. */
int qb_convert(int src_dtype, int dst_dtype, int64_t n, const void *src,
               void *dst, void *stream);

/* complex128 <-> interleaved real embedding E[2i+a, 2j+b] (row-major
 * 2m x 2n): complex QR / SVD run on the real kernels through it.
 * qb_extract_complex: out[i, c] = E[2i, c*col_step] + 1j E[2i+1, c*col_step]. */
int qb_embed_complex(int64_t m, int64_t n, const void *z, void *E, void *stream);
int qb_extract_complex(int64_t m, int64_t ncols, int64_t col_step, const void *E,
                       int64_t ld, void *out, void *stream);

/* ---- decompositions ---------------------------------------------------- */
/*
 * Stabilised thin QR of a row-major contiguous m x n matrix X (m >= n):
 * Q (m x n, row-major), R (n x n, row-major, upper triangular, diag >= 0).
 * Either output may be NULL.  decomp.py:2055-2216.
 */
int qb_qr_stab(int dtype, int64_t m, int64_t n, const void *X, void *Q,
               void *R, int stabilized, void *workspace,
               size_t workspace_bytes, void *stream);
int64_t qb_qr_workspace(int dtype, int64_t m, int64_t n);

/*
 * Thin SVD of a row-major contiguous m x n matrix X by one-sided block
 * Jacobi: U (m x k), S (k, real, descending), VH (k x n), k = min(m, n).
 * sweeps_out / offnorm_out (host) report convergence.
 */
int qb_svd(int dtype, int64_t m, int64_t n, const void *X, void *U, void *S,
           void *VH, void *workspace, size_t workspace_bytes,
           int *sweeps_out, void *stream);
int64_t qb_svd_workspace(int dtype, int64_t m, int64_t n);

/*
 * Truncated SVD with the reference's epilogue fused in: svd_truncated,
 * quimb/tensor/decomp.py:829-898 with the keep rule
 * _compute_number_svals_to_keep (:901-937; cutoff_mode 1..6 = QB_CUTOFF_*,
 * cutoff <= 0 and renorm == 0: only max_bond applies, max_bond = -1: none),
 * _compute_svals_renorm_factor (:940-965), the trim (:968-1029) and
 * _do_absorb (:693-721; `absorb` = QB_ABSORB_*).  X is row-major m x n with
 * m >= n (the host layer passes the transpose of a wide matrix and swaps the
 * roles of the factors).  Only the kept rank k = *n_keep is written, compactly:
 * U as m x k (leading dimension k), S as k values (renormalised), VH as
 * k x n; factors the absorb mode does not request are not touched and may be
 * NULL.  The caller provides room for k = n.  *trunc_error = sqrt(sum of the
 * discarded s^2) (info["error"]); *n_null = number of kept singular values
 * that are exactly zero (their rows of VH are zero: complete them if an
 * isometry is needed).  Workspace: qb_svd_workspace(dtype, m, n) bytes.
 */
/* TEST / debug entry, host only: the sweep schedule of the Jacobi SVD (which
 * column-block pairs rotate in which round of which stream); five int32 per
 * pair: phase, group, round, p, q.  Returns the number of pairs. */
int64_t qb_debug_jacobi_schedule(int nblk, int groups, int32_t *out,
                                 int64_t capacity, int *groups_used);

int qb_svd_trunc(int dtype, int64_t m, int64_t n, const void *X, double cutoff,
                 int cutoff_mode, int64_t max_bond, int absorb, int renorm,
                 void *U, void *S, void *VH, int64_t *n_keep,
                 double *trunc_error, int64_t *n_null, void *workspace,
                 size_t workspace_bytes, int *sweeps_out, void *stream);

/*
 * Host-side truncation rule of svd_truncated's numba core,
 * decomp.py:901-937 (_compute_number_svals_to_keep_numba) followed by the
 * max_bond clamp of decomp.py:1000-1004.  s: host array, descending.
 */
int qb_svals_to_keep(const double *s, int64_t n, double cutoff,
                     int cutoff_mode, int64_t max_bond, int renorm,
                     int64_t *n_keep, double *renorm_factor,
                     double *trunc_error);

/* ---- device queries / microbenchmarks ---------------------------------- */
/* ------------------------------------------------------------------------
 * Peer-memory exchange of the bond-sharded two-site eigensolve (multi-GPU,
 * one process per GPU).  Replaces, on the data path, the ncclAllGather of the
 * Lanczos vector and the ncclAllReduce of the Gram-Schmidt inner products
 * that a sharded TNLinearOperator._matvec / eigsh would issue
 * (quimb/tensor/tensor_core.py:12393-12417, quimb/linalg/scipy_linalg.py:
 * 113-128): ONE kernel each, storing into the peers' HBM over NVLink.
 *
 * Every rank allocates one block (qb_p2p_alloc: cudaMalloc'ed and zeroed by
 * the library -- the only allocation the library owns, released by
 * qb_p2p_free), exports it (64-byte CUDA IPC handle) and imports the blocks
 * of its peers; `peer_bufs[world]` lists the block bases in rank order, the
 * caller's own block at index `rank`.  `epoch` counts the calls of each kind
 * (1, 2, ...; the same sequence on every rank); `scratch` is 8 bytes of
 * zeroed device memory (CTA counter, error word: non-zero after a kernel gave
 * up waiting ~4 s for a peer).
 */
int64_t qb_p2p_block_bytes(int64_t gather_bytes);
int64_t qb_p2p_data_offset(int64_t gather_bytes, int parity);
int qb_p2p_alloc(int64_t bytes, void **ptr);
int qb_p2p_free(void *ptr);
int qb_p2p_export(void *ptr, unsigned char *handle64);
int qb_p2p_import(const unsigned char *handle64, void **peer_ptr);
int qb_p2p_unimport(void *peer_ptr);
/* all ranks: slab `x` (bytes, 16-byte granular) -> byte offset dst_off of the
 * gather buffer of parity (epoch & 1) on EVERY rank; when the kernel has
 * completed the whole vector is resident in the local block. */
int qb_p2p_allgather(void *const *peer_bufs, int world, int rank, const void *x,
                     int64_t bytes, int64_t dst_off, int64_t gather_bytes,
                     uint64_t epoch, void *scratch, void *stream);
/* in-place sum over the ranks of m <= 64 doubles, bit-identical everywhere */
int qb_p2p_allreduce_small(void *const *peer_bufs, int world, int rank, void *x,
                           int m, uint64_t epoch, void *scratch, void *stream);

/* sustained DMMA (fp64 tensor core) rate of the current device in TFLOP/s */
int qb_measure_dmma_peak(double *tflops, void *stream);
/* tuning hook: with QB_TRACE=1 in the environment the contraction kernels stamp
 * %globaltimer at their phase boundaries (entry, tables, prologue issued,
 * first tile landed, main loop done, stream-K fix-up done, stores issued) into
 * a device buffer laid out [cta][4 segments][8 phases]; this copies the first
 * `count` uint64 entries to `host_out`, clears the buffer and returns 0
 * (1 if tracing is off). */
int qb_debug_trace_read(unsigned long long *host_out, int64_t count);

#ifdef __cplusplus
}
#endif
#endif /* QUIMB_B200_H */
